"""CPU oracle for the GatedGCN message-passing hot path.  TEST INFRASTRUCTURE ONLY.

This file is a DGL-free restatement, in plain PyTorch, of the reference forward
(`/root/reference`, commit 102a61d6).  It is the checker the CUDA path is compared
against; nothing in the product package (`gnnome_assembly_b200/`) imports it.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.

Parity pin status: the reference ships no tests, golden vectors or saved activations
 and its own forward needs DGL, which is not installable here.  The pin
we do have: `tests/golden/make_golden.py` imports the UNMODIFIED reference `layers/` and
`models/` packages on top of `oracle/dgl_shim.py` (a restatement of the six DGL calls the
path makes, from DGL's published semantics) and stores inputs/outputs/gradients under
`tests/golden/`; `tests/test_oracle.py` checks this oracle against those fixtures.  So the
oracle is pinned to the reference's own Python code; the DGL primitives underneath are
"parity unpinned" (restated from documentation, no DGL binary to run).

Every function cites the reference file:line it follows.  dtype follows the inputs, so
the same code serves as the fp32 oracle and as the fp64 tie-breaker.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# DGL primitives restated (index_select == gather along edges, index_add_ == sum reducer)
# --------------------------------------------------------------------------------------
def u_add_v(src, dst, a, b):
    """dgl.function.u_add_v + apply_edges (layers/gated_gcn_full.py:120,133)."""
    return a.index_select(0, src) + b.index_select(0, dst)


def u_mul_e_sum(src, dst, num_nodes, node_feat, edge_feat):
    """update_all(fn.u_mul_e, fn.sum): out[v] = sum_{i: dst_i = v} node_feat[src_i] * edge_feat[i]
    (layers/gated_gcn_full.py:128,141).  Zero in-degree rows stay 0 (DGL sum reducer)."""
    out = torch.zeros(num_nodes, node_feat.shape[1], dtype=node_feat.dtype, device=node_feat.device)
    out.index_add_(0, dst, node_feat.index_select(0, src) * edge_feat)
    return out


def copy_e_sum(dst, num_nodes, edge_feat):
    """update_all(fn.copy_e, fn.sum) (layers/gated_gcn_full.py:129,142)."""
    out = torch.zeros(num_nodes, edge_feat.shape[1], dtype=edge_feat.dtype, device=edge_feat.device)
    out.index_add_(0, dst, edge_feat)
    return out


# --------------------------------------------------------------------------------------
# Layer
# --------------------------------------------------------------------------------------
class OracleGatedGCN(nn.Module):
    """layers/gated_gcn_full.py:35-59 (parameters) and :99-157 (forward)."""

    def __init__(self, in_channels, out_channels, batch_norm, dropout=0, residual=True):
        super().__init__()
        self.dropout = dropout
        self.batch_norm = batch_norm
        self.residual = residual and in_channels == out_channels   # :41-42
        for name in ("A_1", "A_2", "A_3", "B_1", "B_2", "B_3"):     # :46-52
            setattr(self, name, nn.Linear(in_channels, out_channels))
        if batch_norm:                                              # :54-56
            self.bn_h = nn.BatchNorm1d(out_channels, track_running_stats=False)
            self.bn_e = nn.BatchNorm1d(out_channels, track_running_stats=False)
        else:                                                       # :57-59
            self.bn_h = nn.LayerNorm(out_channels)
            self.bn_e = nn.LayerNorm(out_channels)

    def forward(self, src, dst, num_nodes, h, e, probe=None):
        """`probe` (test aid, a dict): records the smallest |pre-ReLU value| seen, so a test can pick
        inputs whose ReLU masks cannot flip between fp32 and fp64."""
        h_in, e_in = h, e                                           # :101-102
        A1h, A2h, A3h = self.A_1(h), self.A_2(h), self.A_3(h)       # :107-109
        B1h, B2h, B3e = self.B_1(h), self.B_2(h), self.B_3(e)       # :111-113

        # forward message passing, :120-130
        e_ji = u_add_v(src, dst, B1h, B2h) + B3e                    # :120-121
        e_ji = self.bn_e(e_ji)                                      # :122
        if probe is not None:
            probe["min_abs_pre"] = min(probe.get("min_abs_pre", 1e30), float(e_ji.detach().abs().min())) \
                if e_ji.numel() else probe.get("min_abs_pre", 1e30)
        e_ji = F.relu(e_ji)                                         # :123
        if self.residual:
            e_ji = e_ji + e_in                                      # :124-125
        sigma_f = torch.sigmoid(e_ji)                               # :127
        num_f = u_mul_e_sum(src, dst, num_nodes, A2h, sigma_f)      # :128
        den_f = copy_e_sum(dst, num_nodes, sigma_f)                 # :129
        h_forward = num_f / (den_f + 1e-6)                          # :130

        # reverse graph (dgl.reverse keeps edge ids; src/dst swap), :115,133-143
        rsrc, rdst = dst, src
        e_ik = u_add_v(rsrc, rdst, B2h, B1h) + B3e                  # :133-134
        e_ik = F.relu(self.bn_e(e_ik))                              # :135-136
        if self.residual:
            e_ik = e_ik + e_in                                      # :137-138
        sigma_b = torch.sigmoid(e_ik)                               # :140
        num_b = u_mul_e_sum(rsrc, rdst, num_nodes, A3h, sigma_b)    # :141
        den_b = copy_e_sum(rdst, num_nodes, sigma_b)                # :142
        h_backward = num_b / (den_b + 1e-6)                         # :143

        h = A1h + h_forward + h_backward                            # :145
        h = self.bn_h(h)                                            # :147
        if probe is not None:
            probe["min_abs_pre"] = min(probe.get("min_abs_pre", 1e30), float(h.detach().abs().min()))
        h = F.relu(h)                                               # :149
        if self.residual:
            h = h + h_in                                            # :151-152
        h = F.dropout(h, self.dropout, training=self.training)      # :154
        return h, e_ji                                              # :155-157

    # independent formulation from the reference's dead UDF code (:61-97): dense mailbox per node
    def forward_mailbox(self, src, dst, num_nodes, h, e):
        A1h, A2h, A3h = self.A_1(h), self.A_2(h), self.A_3(h)
        B1h, B2h, B3e = self.B_1(h), self.B_2(h), self.B_3(e)
        e_new = F.relu(self.bn_e(B1h[src] + B2h[dst] + B3e))        # bn applied unconditionally as in :122
        if self.residual:
            e_new = e_new + e
        sig = torch.sigmoid(e_new)
        hf = torch.zeros_like(A1h)
        hb = torch.zeros_like(A1h)
        for v in range(num_nodes):
            m = (dst == v).nonzero().flatten()
            if m.numel():                                           # reduce_forward :72-78
                hf[v] = (sig[m] * A2h[src[m]]).sum(0) / (sig[m].sum(0) + 1e-6)
            m = (src == v).nonzero().flatten()
            if m.numel():                                           # reduce_backward :91-97
                hb[v] = (sig[m] * A3h[dst[m]]).sum(0) / (sig[m].sum(0) + 1e-6)
        hn = F.relu(self.bn_h(A1h + hf + hb))
        if self.residual:
            hn = hn + h
        return hn, e_new


class OracleProcessor(nn.Module):
    """layers/processor.py:8-20."""

    def __init__(self, num_layers, hidden_features, batch_norm):
        super().__init__()
        self.convs = nn.ModuleList(
            [OracleGatedGCN(hidden_features, hidden_features, batch_norm) for _ in range(num_layers)])

    def forward(self, src, dst, num_nodes, h, e):
        for conv in self.convs:
            h, e = conv(src, dst, num_nodes, h, e)
        return h, e


class OracleScorePredictor(nn.Module):
    """layers/score_predictor.py:6-25."""

    def __init__(self, in_features, hidden_edge_scores):
        super().__init__()
        self.W1 = nn.Linear(3 * in_features, hidden_edge_scores)
        self.W2 = nn.Linear(hidden_edge_scores, 1)

    def forward(self, src, dst, x, e):
        data = torch.cat((x.index_select(0, src), x.index_select(0, dst), e), dim=1)   # :13
        return self.W2(torch.relu(self.W1(data)))                                      # :15-17


class OracleModel(nn.Module):
    """models/full_graph.py:11-29.  Same constructor arguments and the same state_dict keys as the
    reference `GraphGatedGCNModel`, so the shipped checkpoints load with strict=True."""

    def __init__(self, node_features, edge_features, hidden_features, hidden_edge_features, num_layers,
                 hidden_edge_scores, batch_norm, nb_pos_enc):
        super().__init__()
        self.linear_pe = nn.Linear(nb_pos_enc + 2, hidden_features)                    # :15
        self.linear1_edge = nn.Linear(edge_features, hidden_edge_features)             # :17
        self.linear2_edge = nn.Linear(hidden_edge_features, hidden_features)           # :18
        self.gnn = OracleProcessor(num_layers, hidden_features, batch_norm)            # :19
        self.predictor = OracleScorePredictor(hidden_features, hidden_edge_scores)     # :20

    def forward(self, src, dst, num_nodes, e, pe):
        x = self.linear_pe(pe)                                                         # :23 (arg x ignored)
        e = self.linear2_edge(torch.relu(self.linear1_edge(e)))                        # :24-26
        x, e = self.gnn(src, dst, num_nodes, x, e)                                     # :27
        return self.predictor(src, dst, x, e)                                          # :28


def bce_loss(scores, y, pos_weight):
    """train.py:209-211,253-255: BCEWithLogitsLoss(pos_weight) on scores.squeeze(-1)."""
    pw = torch.as_tensor([pos_weight], dtype=scores.dtype, device=scores.device)
    return F.binary_cross_entropy_with_logits(scores.squeeze(-1), y.to(scores.dtype), pos_weight=pw)


def rel_err(a, b):
    """max|a-b| / max|b| — the parity figure BASELINE.json's 1e-4 tolerance is stated in."""
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def grads_close(got, ref, rtol, atol_frac=1e-6):
    """Compare two {name: grad} dicts.  Under BatchNorm the biases of A_1, B_1, B_2, B_3 have an
    exactly-zero true gradient (mean subtraction removes them), so what autograd
    returns there is rounding noise: errors are therefore judged against rtol*max|ref_k| plus an
    absolute floor of atol_frac * (largest gradient entry in the whole model).  Returns the list
    of offending (name, err, bound)."""
    scale = max(float(v.abs().max()) for v in ref.values())
    bad = []
    for k, r in ref.items():
        err = float((got[k].detach().double().cpu() - r.detach().double().cpu()).abs().max())
        bound = rtol * float(r.abs().max()) + atol_frac * scale
        if not err <= bound:
            bad.append((k, err, bound))
    return bad
