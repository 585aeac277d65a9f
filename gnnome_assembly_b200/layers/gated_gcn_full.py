"""GatedGCN_1d — drop-in for the reference layer (layers/gated_gcn_full.py:10-157).

Same constructor, same parameter names / shapes (A_1..A_3, B_1..B_3 as nn.Linear, bn_h / bn_e as
BatchNorm1d(track_running_stats=False) or LayerNorm), same `forward(g, h, e) -> (h, e)` in the caller's
node / edge-id order.  The arithmetic runs in the sm_100a kernels behind the C ABI; the torch modules
here are only parameter containers.  Unlike the reference the forward does NOT write intermediate
tensors into `g.ndata` / `g.edata` (nothing reads them).
"""
import torch
import torch.nn as nn

from .. import functional as GF
from ..plan import plan_for


class GatedGCN_1d(nn.Module):
    def __init__(self, in_channels, out_channels, batch_norm, dropout=0, residual=True):
        super().__init__()
        self.dropout = dropout
        self.batch_norm = batch_norm
        self.residual = residual
        if in_channels != out_channels:                      # gated_gcn_full.py:41-42
            self.residual = False
        if in_channels != out_channels or out_channels not in (64, 128, 256):
            raise NotImplementedError(
                "gnnome_assembly_b200 kernels are built for in_channels == out_channels in {64, 128, 256}")
        dtype = torch.float32
        self.A_1 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.A_2 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.A_3 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.B_1 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.B_2 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.B_3 = nn.Linear(in_channels, out_channels, dtype=dtype)
        if batch_norm:
            self.bn_h = nn.BatchNorm1d(out_channels, track_running_stats=False)
            self.bn_e = nn.BatchNorm1d(out_channels, track_running_stats=False)
        else:
            self.bn_h = nn.LayerNorm(out_channels)
            self.bn_e = nn.LayerNorm(out_channels)

    def forward_internal(self, plan, h, e, arena=None):
        """h [N,d], e [E,d] in the plan's internal node / edge order.  arena: the pass's flat.GradArena (model seam)."""
        if self.dropout and self.training:
            raise NotImplementedError("dropout > 0 is never used on the reference path (processor.py:12)")
        norm = GF.NORM_BATCH if self.batch_norm else GF.NORM_LAYER
        return GF.gated_gcn_layer(plan, norm, self.residual, h, e, self, arena)

    def forward(self, g, h, e):
        from ..flat import ensure_flat, packed_node_weights
        if packed_node_weights(self) is None:              # stand-alone layer: its own flat parameter buffer
            ensure_flat(self)
        plan = plan_for(g, h.device)
        e_int = GF.permute_rows(e, plan.perm, plan.inv_perm)
        h, e_int = self.forward_internal(plan, GF.permute_rows(h, plan.node_perm, plan.node_inv), e_int)
        return (GF.permute_rows(h, plan.node_inv, plan.node_perm),
                GF.permute_rows(e_int, plan.inv_perm, plan.perm))
