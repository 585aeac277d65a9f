"""Parity at the sizes the headline numbers are measured at (pytest -m gpu, through the C ABI).

Round-1 gradient checks stopped at E = 2.5 k (one 128-row tile per persistent CTA).  These tests run the
tcgen05 path where every persistent CTA loops over ~20 tiles (mbarrier phase flips, TMEM accumulator
swaps, split-K weight gradients in one wave, W-resident B) and hold ALL gradients to the CPU oracle:

  * BASELINE.json configs[1] graph (chr19-like, E ~ 373 k), d = 128, L = 2, BatchNorm: logits, loss and every
    parameter gradient vs the fp32 CPU oracle (layers/gated_gcn_full.py:99-157 restated);
  * one layer d = 256 at E ~ 100 k vs the fp64 oracle, inputs + parameters;
  * configs[4] point: 1M-edge graph, d = 256, L = 1 forward + backward vs the oracle;
  * configs[2]: shipped model_15xchr19.pt (L = 16, d = 256) on the FULL chr21-like graph, max and median
    relative logit error <= 1e-4 (BASELINE.json's bar);
  * tcgen05 path vs the exact-fp32 FFMA path of the same library on the full training step (loss trajectory).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4          # BASELINE.json: edge logits <= 1e-4 relative fp32


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _graph_tensors(g):
    src = torch.from_numpy(g.src.astype(np.int64))
    dst = torch.from_numpy(g.dst.astype(np.int64))
    return src, dst, torch.from_numpy(g.e), torch.from_numpy(g.pe), torch.from_numpy(g.y)


def _model_pair(d, L, dev, seed=0):
    import gnnome_assembly_b200 as gg
    from oracle.gatedgcn_oracle import OracleModel
    torch.manual_seed(seed)
    oracle = OracleModel(1, 2, d, 16, L, 64, True, 16)
    with torch.no_grad():                                   # move the norm affines off their 1 / 0 init
        for n_, p in oracle.named_parameters():
            if "bn_" in n_:
                p.add_(0.2 * torch.randn_like(p))
    model = gg.GraphGatedGCNModel(1, 2, d, 16, L, 64, True, 16)
    model.load_state_dict(oracle.state_dict(), strict=True)
    return model.to(dev), oracle


def _fwd_bwd_both(model, oracle, g, dev):
    import gnnome_assembly_b200 as gg
    from oracle.gatedgcn_oracle import bce_loss
    src, dst, e, pe, y = _graph_tensors(g)
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    s = model(graph, None, e.to(dev), pe.to(dev))
    loss = bce_loss(s, y.to(dev), 1 / 16.5)
    loss.backward()
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    r = oracle(src, dst, g.num_nodes, e, pe)
    rloss = bce_loss(r, y, 1 / 16.5)
    rloss.backward()
    return s, loss, r, rloss


@pytest.fixture(scope="module")
def chr19_graph():
    from gnnome_assembly_b200.synth import make_assembly_graph
    return make_assembly_graph("chr19", seed=0)


def test_chr19_size_gradients_tc_path(chr19_graph):
    """configs[1] graph, d=128, L=2, BN: 2,911 row tiles over 148 persistent CTAs -> the fused BnBwdATx GEMM,
    gemm_edge_gate and the one-wave split-K weight gradients all run their multi-tile loops under a gradient check."""
    dev = _dev()
    from gnnome_assembly_b200 import _lib
    from oracle.gatedgcn_oracle import grads_close, rel_err
    assert _lib.set_tc_mode(1) in (0, 1)
    g = chr19_graph
    assert g.num_edges > 148 * 128 * 8
    model, oracle = _model_pair(128, 2, dev)
    s, loss, r, rloss = _fwd_bwd_both(model, oracle, g, dev)
    assert rel_err(s, r) < TOL
    assert abs(float(loss.detach()) - float(rloss.detach())) < 1e-5 * max(1.0, abs(float(rloss.detach())))
    bad = grads_close({k: p.grad for k, p in model.named_parameters()},
                      {k: p.grad for k, p in oracle.named_parameters()}, rtol=2e-3, atol_frac=1e-5)
    assert bad == [], bad[:4]


@pytest.mark.parametrize("d", [128, 256])
def test_layer_gradients_100k_edges_vs_fp64(d):
    """One GatedGCN layer at E ~ 100 k (780 row tiles: > 5 per CTA) against the fp64 oracle: outputs, input
    gradients and every parameter gradient, tcgen05 path."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200.synth import make_assembly_graph
    from oracle import gatedgcn_oracle as O
    g = make_assembly_graph("chr19", seed=3, target_edges=100_000, pe_dim=0)
    src = torch.from_numpy(g.src.astype(np.int64))
    dst = torch.from_numpy(g.dst.astype(np.int64))
    torch.manual_seed(d)
    ref = O.OracleGatedGCN(d, d, True)
    with torch.no_grad():
        for n_, p in ref.named_parameters():
            if "bn_" in n_:
                p.add_(0.3 * torch.randn_like(p))
    ours = gg.layers.GatedGCN_1d(d, d, True)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.to(dev)
    h, e = torch.randn(g.num_nodes, d), torch.randn(g.num_edges, d)
    gh, ge = torch.randn(g.num_nodes, d), torch.randn(g.num_edges, d)
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    hd, ed = h.to(dev).requires_grad_(), e.to(dev).requires_grad_()
    ho, eo = ours(graph, hd, ed)
    ((ho * gh.to(dev)).sum() + (eo * ge.to(dev)).sum()).backward()
    ref = ref.double()
    h64, e64 = h.double().requires_grad_(), e.double().requires_grad_()
    ho64, eo64 = ref(src, dst, g.num_nodes, h64, e64)
    ((ho64 * gh.double()).sum() + (eo64 * ge.double()).sum()).backward()
    assert O.rel_err(ho, ho64) < 2e-5 and O.rel_err(eo, eo64) < 2e-5
    gr = {k: p.grad for k, p in ours.named_parameters()}
    gr64 = {k: p.grad for k, p in ref.named_parameters()}
    gr["__h"], gr["__e"], gr64["__h"], gr64["__e"] = hd.grad, ed.grad, h64.grad, e64.grad
    # 100 k x d pre-ReLU values with unit-variance inputs: ~10 of them sit within fp32 rounding of 0 and flip their ReLU
    # mask against the fp64 oracle; each flip moves one row of dB3 (and one entry of the node sums) by |g_eo e_in| = O(1..10),
    # i.e. a few 1e-3 of the largest entry.  The max-norm bound therefore is 1e-2, and the discriminating check is the
    # Frobenius one: sparse flips stay below 1e-3 of the tensor's norm, whereas ONE wrong / missing 128-row tile out of
    # 780 (a broken phase flip or accumulator swap) would show up as ~3e-2.
    # (a flipped element moves its own row of the INPUT gradient __e by |g_eo gamma rstd B3| = O(0.1), comparable to that
    # tensor's largest entry: the input gradients are judged by the Frobenius criterion only)
    params = [k for k in gr64 if not k.startswith("__")]
    bad = O.grads_close({k: gr[k] for k in params}, {k: gr64[k] for k in params}, rtol=1e-2, atol_frac=1e-5)
    assert bad == [], bad[:4]
    scale = max(float(v.norm()) for v in gr64.values())
    for k, r in gr64.items():
        err = float((gr[k].detach().double().cpu() - r).norm())
        assert err <= 2e-3 * float(r.norm()) + 1e-6 * scale, (k, err, float(r.norm()))


def test_config5_point_1m_edges_d256():
    """configs[4] sweep point (1M edges, d=256, L=1, fwd+bwd): timing-only in round 1, now held to the oracle."""
    dev = _dev()
    from gnnome_assembly_b200.synth import make_assembly_graph
    from oracle.gatedgcn_oracle import grads_close, rel_err
    g = make_assembly_graph("chr19", seed=0, target_edges=1_000_000)
    assert g.num_edges > 900_000
    model, oracle = _model_pair(256, 1, dev, seed=5)
    s, loss, r, rloss = _fwd_bwd_both(model, oracle, g, dev)
    assert rel_err(s, r) < TOL
    assert abs(float(loss.detach()) - float(rloss.detach())) < 1e-5 * max(1.0, abs(float(rloss.detach())))
    bad = grads_close({k: p.grad for k, p in model.named_parameters()},
                      {k: p.grad for k, p in oracle.named_parameters()}, rtol=2e-3, atol_frac=1e-5)
    assert bad == [], bad[:4]


def test_config3_full_chr21_shipped_checkpoint(ckpt_path):
    """configs[2]: the shipped model_15xchr19.pt (L=16, d=256) on the full chr21-like graph; edge logits within
    1e-4 of the CPU oracle, max-relative (BASELINE.json's figure) and median-relative."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200.synth import make_assembly_graph
    from oracle.gatedgcn_oracle import OracleModel, rel_err
    g = make_assembly_graph("chr21", seed=0)
    sd = torch.load(ckpt_path, map_location="cpu")
    model = gg.GraphGatedGCNModel(1, 2, 256, 16, 16, 64, True, 16)
    model.load_state_dict(sd, strict=True)
    model.eval().to(dev)
    oracle = OracleModel(1, 2, 256, 16, 16, 64, True, 16)
    oracle.load_state_dict(sd, strict=True)
    src, dst, e, pe, _ = _graph_tensors(g)
    with torch.no_grad():
        s = model(gg.AssemblyGraph(src, dst, g.num_nodes), None, e.to(dev), pe.to(dev)).cpu()
        r = oracle(src, dst, g.num_nodes, e, pe)
    assert rel_err(s, r) < TOL
    med = ((s - r).abs() / r.abs().clamp_min(1e-3)).median()
    assert float(med) < TOL


def test_training_trajectory_tc_equals_ffma(chr19_graph):
    """Four Adam steps of the bench workload (d=128, L=4 here) on the tcgen05 path and on the exact-fp32 FFMA path
    of the same library: the loss trajectories agree to 1e-4 relative, i.e. the gradients that produce the headline
    number drive the optimiser the same way the true-fp32 ones do."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200 import _lib
    from oracle.gatedgcn_oracle import bce_loss
    g = chr19_graph
    src, dst, e, pe, y = _graph_tensors(g)
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    e, pe, y = e.to(dev), pe.to(dev), y.to(dev)

    def run(mode):
        old = _lib.set_tc_mode(mode)
        try:
            torch.manual_seed(0)
            model = gg.GraphGatedGCNModel(1, 2, 128, 16, 4, 64, True, 16).to(dev)
            opt = torch.optim.Adam(model.parameters(), lr=1e-3)
            out = []
            for _ in range(4):
                loss = bce_loss(model(graph, None, e, pe), y, 1 / 16.5)
                opt.zero_grad()
                loss.backward()
                opt.step()
                out.append(float(loss.detach()))
            return out
        finally:
            _lib.set_tc_mode(old)

    a, b = run(1), run(0)
    assert all(np.isfinite(a)) and a[-1] < a[0]
    for x, y_ in zip(a, b):
        assert abs(x - y_) < 1e-4 * max(1.0, abs(y_)), (a, b)


@pytest.mark.parametrize("d,L,bn", [(64, 2, True), (128, 3, True), (128, 2, False), (256, 2, True)])
def test_one_call_model_path_equals_per_op_path(d, L, bn):
    """gg_model_fwd / gg_model_bwd (one C-ABI call per direction, gradients into the flat arena) against the per-op
    bindings (one autograd node per encoder / layer / predictor): same kernels in the same order, so logits and every
    gradient agree to fp64-atomic reordering noise; and both against the CPU oracle."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200.synth import make_assembly_graph
    from oracle.gatedgcn_oracle import OracleModel, bce_loss, grads_close, rel_err
    g = make_assembly_graph("chr19", seed=11, genome_len=2_000_000)
    src, dst, e, pe, y = _graph_tensors(g)
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    torch.manual_seed(d + L)
    oracle = OracleModel(1, 2, d, 16, L, 64, bn, 16)
    outs = {}
    for per_op in (False, True):
        model = gg.GraphGatedGCNModel(1, 2, d, 16, L, 64, bn, 16)
        model.load_state_dict(oracle.state_dict(), strict=True)
        model.to(dev)
        model.per_op_path = per_op
        s = model(graph, None, e.to(dev), pe.to(dev))
        bce_loss(s, y.to(dev), 1 / 16.5).backward()
        outs[per_op] = (s.detach(), {k: p.grad.clone() for k, p in model.named_parameters()})
        with torch.no_grad():
            s_inf = model(graph, None, e.to(dev), pe.to(dev))              # inference layout of the workspace
        assert rel_err(s_inf, s) < 1e-6
    assert rel_err(outs[False][0], outs[True][0]) < 1e-6
    assert grads_close(outs[False][1], outs[True][1], rtol=1e-4, atol_frac=1e-5) == []   # (split-K fp32 atomics: order varies run to run)
    # the fp64 oracle as the reference: on a default-initialised model several gradients (the last layer's B_1 / B_2) are
    # tiny differences of large sums, and the fp32 oracle's own summation noise there is as large as ours
    oracle = oracle.double()
    r = oracle(src, dst, g.num_nodes, e.double(), pe.double())
    bce_loss(r, y.double(), 1 / 16.5).backward()
    assert rel_err(outs[False][0], r) < TOL
    bad = grads_close(outs[False][1], {k: p.grad for k, p in oracle.named_parameters()}, rtol=2e-3, atol_frac=1e-4)
    assert bad == [], bad


def test_gradients_live_in_the_arena_and_survive_a_second_pass():
    """flat.GradArena: after backward every .grad IS its slot of the pass's arena (no copies); a second forward/backward
    gets a fresh arena, so gradients of the first pass are not overwritten behind the caller's back."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200.synth import make_assembly_graph
    from oracle.gatedgcn_oracle import bce_loss
    g = make_assembly_graph("chr19", seed=5, genome_len=1_000_000)
    src, dst, e, pe, y = _graph_tensors(g)
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    torch.manual_seed(0)
    model = gg.GraphGatedGCNModel(1, 2, 128, 16, 2, 64, True, 16).to(dev)
    for per_op in (False, True):
        model.per_op_path = per_op
        model.zero_grad(set_to_none=True)
        bce_loss(model(graph, None, e.to(dev), pe.to(dev)), y.to(dev), 1 / 16.5).backward()
        arena = model._gg_last_arena
        base = arena.tensor().data_ptr()
        for p, off in arena.layout.entries:
            assert p.grad.data_ptr() == base + 4 * off, "a gradient was copied out of the arena"
        first = {k: p.grad.clone() for k, p in model.named_parameters()}
        kept = {k: p.grad for k, p in model.named_parameters()}
        model.zero_grad(set_to_none=True)
        bce_loss(model(graph, None, (2 * e).to(dev), pe.to(dev)), y.to(dev), 1 / 16.5).backward()
        assert model._gg_last_arena is not arena
        for k in first:
            assert torch.equal(first[k], kept[k])


def test_model_abi_with_a_gapped_offset_table():
    """include/gnnome_b200.h, gg_model_bwd: a table whose only gaps are alignment padding gets ONE memset of the whole
    gradient span at phase 0; any other table (here: 64 floats between consecutive tensors) makes every op zero its own
    outputs.  Both must give the same gradients, and the gapped call must not touch the gaps."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200 import _lib, flat as gflat, functional as GF
    from gnnome_assembly_b200.synth import make_assembly_graph
    g = make_assembly_graph("chr19", seed=3, genome_len=1_500_000)
    src, dst, e, pe, y = _graph_tensors(g)
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    plan = gg.plan_for(graph, dev)
    torch.manual_seed(1)
    model = gg.GraphGatedGCNModel(1, 2, 128, 16, 2, 64, True, 16).to(dev)
    layout = gflat.ensure_flat(model)
    flatbuf = model.__dict__["_gg_flat"][1]
    desc, offs, n = GF.model_call_table(model, layout)
    C, lib = _lib.C, _lib.lib()
    E = plan.num_edges
    e_d, pe_d = e.to(dev), pe.to(dev)
    g_scores = torch.randn(E, 1, device=dev) / E
    st = torch.cuda.current_stream(dev).cuda_stream

    def run(params, table, grads):
        ws = torch.empty(lib.gg_model_workspace_floats(plan.handle, C.byref(desc), 0), device=dev)
        bws = torch.empty(lib.gg_model_workspace_floats(plan.handle, C.byref(desc), 2), device=dev)
        scores = torch.empty(E, 1, device=dev)
        tab = (C.c_int64 * n)(*table)
        GF.check(lib.gg_model_fwd(plan.handle, C.byref(desc), GF.ptr(params), tab, n, GF.ptr(e_d), GF.ptr(pe_d), 1,
                                  GF.ptr(ws), GF.ptr(scores), st), "gg_model_fwd")
        GF.check(lib.gg_model_bwd(plan.handle, C.byref(desc), GF.ptr(params), tab, n, GF.ptr(g_scores), GF.ptr(ws),
                                  GF.ptr(bws), GF.ptr(grads), 0, desc.layers + 2, st, None), "gg_model_bwd")
        torch.cuda.synchronize()
        return scores

    packed = [int(offs[i]) for i in range(n)]
    d, he, H = desc.d, desc.hidden_edge, desc.hidden_score
    sizes = [d * desc.node_in, d, he * desc.edge_in, he, d * he, d, H * 3 * d, H, H, 1]        # the header's table
    sizes += [5 * d * d, 5 * d, d * d, d, d, d, d, d] * desc.layers
    gap, gapped, cur = 64, [0] * n, 64
    for i in sorted(range(n), key=lambda i: packed[i]):
        gapped[i] = cur
        cur += (sizes[i] + 3) // 4 * 4 + gap                    # every tensor on a 16-byte boundary (header contract)
    flat2 = torch.zeros(cur, device=dev)
    for i in range(n):
        flat2[gapped[i]:gapped[i] + sizes[i]] = flatbuf[packed[i]:packed[i] + sizes[i]]
    g1 = torch.full((layout.total,), 7.0, device=dev)
    g2 = torch.full((cur,), 7.0, device=dev)
    s1 = run(flatbuf, packed, g1)
    s2 = run(flat2, gapped, g2)
    assert torch.equal(s1, s2)
    scale = float(g1.abs().max())
    covered = torch.zeros(cur, dtype=torch.bool, device=dev)
    for i in range(n):
        a, b = g1[packed[i]:packed[i] + sizes[i]], g2[gapped[i]:gapped[i] + sizes[i]]
        assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max()) + 1e-6 * scale, i   # split-K atomics reorder
        covered[gapped[i]:gapped[i] + sizes[i]] = True
    assert bool((g2[~covered] == 7.0).all())                   # the gaps belong to the caller
    # a table that breaks the alignment contract is refused, not executed
    bad = list(gapped)
    bad[1] += 1
    with pytest.raises(RuntimeError, match="multiple of 4"):
        run(flat2, bad, g2)
