"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through the
C ABI (ctypes) and compares with the CPU oracle / the golden vectors made by the reference's own code.

Tolerances: BASELINE.json states <= 1e-4 relative fp32 on the edge logits (max|a-b| / max|b|).
Gradients: rtol 2e-3 of each tensor's max entry plus a floor of 1e-5 x the largest gradient in the model
(the biases that BatchNorm cancels have a true gradient of exactly 0, see oracle.grads_close)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _oracle():
    from oracle import gatedgcn_oracle
    return gatedgcn_oracle


# ------------------------------------------------------------------------------------------ plan
@pytest.mark.parametrize("n,m", [(1, 0), (5, 1), (96, 700), (1000, 9000)])
def test_plan_arrays(n, m):
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan
    rng = np.random.default_rng(n + m)
    src = rng.integers(0, n, m).astype(np.int32)
    dst = rng.integers(0, n, m).astype(np.int32)
    plan = GraphPlan(torch.from_numpy(src), torch.from_numpy(dst), n, dev, relabel=False)
    assert np.array_equal(plan.node_perm.cpu().numpy(), np.arange(n))
    perm = plan.perm.cpu().numpy()
    assert np.array_equal(perm, np.argsort(dst, kind="stable"))
    assert np.array_equal(plan.inv_perm.cpu().numpy()[perm], np.arange(m))
    assert np.array_equal(plan.src.cpu().numpy(), src[perm])
    assert np.array_equal(plan.dst.cpu().numpy(), dst[perm])
    in_ptr = plan.array("in_ptr").cpu().numpy()
    assert np.array_equal(np.diff(in_ptr), np.bincount(dst, minlength=n))
    out_ptr = plan.array("out_ptr").cpu().numpy()
    out_eid = plan.array("out_eid").cpu().numpy()
    assert np.array_equal(np.diff(out_ptr), np.bincount(src, minlength=n))
    assert np.array_equal(np.sort(out_eid), np.arange(m))
    isrc = src[perm]
    for u in range(min(n, 50)):
        seg = out_eid[out_ptr[u]:out_ptr[u + 1]]
        assert np.all(isrc[seg] == u) and np.all(np.diff(seg) > 0)


@pytest.mark.parametrize("n,m", [(7, 0), (96, 700), (5000, 40000)])
def test_plan_relabelled(n, m):
    """Breadth-first node relabelling: a permutation of the nodes; every array is consistent with it."""
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan
    rng = np.random.default_rng(n * 7 + m)
    src = rng.integers(0, n, m).astype(np.int32)
    dst = rng.integers(0, n, m).astype(np.int32)
    plan = GraphPlan(torch.from_numpy(src), torch.from_numpy(dst), n, dev)
    node_perm = plan.node_perm.cpu().numpy()
    node_inv = plan.node_inv.cpu().numpy()
    assert np.array_equal(np.sort(node_perm), np.arange(n))
    assert np.array_equal(node_inv[node_perm], np.arange(n))
    perm = plan.perm.cpu().numpy()
    assert np.array_equal(np.sort(perm), np.arange(m))
    isrc, idst = plan.src.cpu().numpy(), plan.dst.cpu().numpy()
    assert np.array_equal(isrc, node_inv[src[perm]]) and np.array_equal(idst, node_inv[dst[perm]])
    assert np.all(np.diff(idst) >= 0)
    in_ptr = plan.array("in_ptr").cpu().numpy()
    assert np.array_equal(np.diff(in_ptr), np.bincount(idst, minlength=n))
    out_ptr, out_eid = plan.array("out_ptr").cpu().numpy(), plan.array("out_eid").cpu().numpy()
    assert np.array_equal(np.diff(out_ptr), np.bincount(isrc, minlength=n))
    assert np.array_equal(np.sort(out_eid), np.arange(m))
    assert np.array_equal(isrc[out_eid], np.repeat(np.arange(n), np.diff(out_ptr)))


@pytest.mark.parametrize("host_build", [False, True], ids=["device", "host"])
def test_plan_relabel_gives_locality(host_build):
    """On an assembly graph (random read ids) the relabelled order puts neighbours close together: the exact
    breadth-first order of the host builder within tens of rows, the two-level region order of the device builder
    (csrc/gg_plan_device.cu) within a few regions of 64 nodes."""
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan
    from gnnome_assembly_b200.synth import make_assembly_graph
    g = make_assembly_graph("chr19", seed=2, genome_len=3_000_000)
    plan = GraphPlan(torch.from_numpy(g.src), torch.from_numpy(g.dst), g.num_nodes, dev, host_build=host_build)
    isrc, idst = plan.src.cpu().numpy().astype(np.int64), plan.dst.cpu().numpy().astype(np.int64)
    y = g.y[plan.perm.cpu().numpy()] > 0                                  # true overlaps only (no random repeats)
    gap = np.abs(isrc - idst)[y]
    assert np.median(gap) < (64 if host_build else 192), np.median(gap)
    assert np.percentile(gap, 95) < 1024, np.percentile(gap, 95)
    assert np.median(np.abs(g.src.astype(np.int64) - g.dst)[g.y > 0]) > 500


@pytest.mark.parametrize("n,m", [(1, 0), (64, 0), (130, 129), (3000, 25000), (50000, 300000)])
def test_device_plan_equals_host_plan(n, m):
    """The device builder (radix sorts) against the round-1 host builder (counting sorts): every array bit for bit
    without relabelling; with relabelling both are valid plans of the same graph (node orders differ by design) and the
    device one is reproducible run to run."""
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan
    rng = np.random.default_rng(n + 3 * m)
    src = torch.from_numpy(rng.integers(0, n, m).astype(np.int32))
    dst = torch.from_numpy(rng.integers(0, max(1, n - n // 8), m).astype(np.int32))      # some nodes without in-edges
    names = ["perm", "inv_perm", "src", "dst", "in_ptr", "out_ptr", "out_eid", "node_perm", "node_inv"]
    a = GraphPlan(src, dst, n, dev, relabel=False)
    b = GraphPlan(src, dst, n, dev, relabel=False, host_build=True)
    for k in names:
        assert torch.equal(a.array(k), b.array(k)), k
    c = GraphPlan(src.to(dev), dst.to(dev), n, dev)                       # device-resident edge list, relabelled
    d = GraphPlan(src, dst, n, dev)
    for k in names:
        assert torch.equal(c.array(k), d.array(k)), k                     # deterministic, host or device input
    node_inv = c.node_inv.cpu().numpy()
    perm = c.perm.cpu().numpy()
    assert np.array_equal(np.sort(c.node_perm.cpu().numpy()), np.arange(n))
    assert np.array_equal(c.src.cpu().numpy(), node_inv[src.numpy()[perm]])
    assert np.array_equal(c.dst.cpu().numpy(), node_inv[dst.numpy()[perm]])
    assert np.all(np.diff(c.dst.cpu().numpy()) >= 0)


def test_plan_rejects_bad_index():
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan
    with pytest.raises(RuntimeError, match="out of range"):
        GraphPlan(torch.tensor([0, 7]), torch.tensor([1, 2]), 5, dev)


# ------------------------------------------------------------------------------------------ linear
@pytest.fixture(params=[1, 0], ids=["tc", "ffma"])
def tc_mode(request):
    """Run a test on the tcgen05 3xTF32 GEMM path (default) and on the FFMA path."""
    from gnnome_assembly_b200 import _lib
    old = _lib.set_tc_mode(request.param)
    yield request.param
    _lib.set_tc_mode(old)


@pytest.mark.parametrize("M,N,K", [(1, 16, 4), (130, 128, 20), (257, 64, 16), (1000, 640, 128), (300, 20, 256),
                                    (5000, 256, 256), (129, 128, 64), (128, 128, 32), (40000, 128, 128),
                                    (33333, 256, 1280), (777, 1280, 256)])
def test_linear_fwd_bwd(M, N, K, tc_mode):
    dev = _dev()
    from gnnome_assembly_b200 import functional as GF
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev, requires_grad=True)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).requires_grad_()
    b = torch.randn(N, device=dev, requires_grad=True)
    y = GF.linear(x, W, b)
    g = torch.randn_like(y)
    y.backward(g)
    x2, W2, b2 = (t.detach().double().requires_grad_() for t in (x, W, b))
    y2 = torch.nn.functional.linear(x2, W2, b2)
    y2.backward(g.double())
    rel = _oracle().rel_err
    # FFMA path: fp32 rounding only.  tcgen05 3xTF32 path: ~2^-20 per product plus the tensor core's
    # truncating fp32 accumulation, which grows with the number of K steps per accumulator.
    tol = 6e-5 if tc_mode else 1e-5
    assert rel(y, y2) < tol
    assert rel(x.grad, x2.grad) < tol
    assert rel(W.grad, W2.grad) < 2 * tol
    assert rel(b.grad, b2.grad) < 2e-5


def test_linear_unpadded_k():
    """K = 18 (linear_pe) and K = 2 (linear1_edge) go through the zero-padding path."""
    dev = _dev()
    from gnnome_assembly_b200 import functional as GF
    torch.manual_seed(0)
    for K, N in ((18, 128), (2, 16)):
        x = torch.randn(333, K, device=dev)
        lin = torch.nn.Linear(K, N).to(dev)
        y = GF.linear(x, lin.weight, lin.bias)
        y.sum().backward()
        ref = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
        assert _oracle().rel_err(y, ref) < 1e-5
        assert lin.weight.grad.shape == (N, K)
        assert _oracle().rel_err(lin.weight.grad, x.double().sum(0).expand(N, K)) < 1e-5


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("E", [0, 1, 17, 4097, 100_003])
def test_edge_encoder_forward_one_pass(d, E):
    """gg_edge_mlp_fwd (models/full_graph.py:24-26 in one kernel) against the same two layers in fp64: the hidden
    activations it keeps for the backward and the encoder output; odd row counts exercise the two-rows-per-warp tail."""
    dev = _dev()
    from gnnome_assembly_b200 import _lib
    from gnnome_assembly_b200._lib import check, ptr
    torch.manual_seed(E + d)
    l1, l2 = torch.nn.Linear(2, 16).to(dev), torch.nn.Linear(16, d).to(dev)
    e = torch.randn(E, 2, device=dev)
    e4 = torch.zeros(E, 4, device=dev)
    e4[:, :2] = e
    W1 = torch.zeros(16, 4, device=dev)
    W1[:, :2] = l1.weight.detach()
    hid = torch.full((E, 16), -7.0, device=dev)
    out = torch.full((E, d), -7.0, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().gg_edge_mlp_fwd(E, d, 16, 4, ptr(e4), ptr(W1), ptr(l1.bias.detach()), ptr(l2.weight.detach()),
                                     ptr(l2.bias.detach()), ptr(hid), ptr(out), st), "gg_edge_mlp_fwd")
    torch.cuda.synchronize()
    if E == 0:
        return
    h64 = torch.relu(torch.nn.functional.linear(e.double(), l1.weight.double(), l1.bias.double()))
    o64 = torch.nn.functional.linear(h64, l2.weight.double(), l2.bias.double())
    assert _oracle().rel_err(hid, h64) < 1e-6
    assert _oracle().rel_err(out, o64) < 1e-6


def test_edge_encoder_forward_other_shapes_are_refused():
    dev = _dev()
    from gnnome_assembly_b200 import _lib
    from gnnome_assembly_b200._lib import ptr
    x = torch.zeros(64, 256, device=dev)
    rc = _lib.lib().gg_edge_mlp_fwd(4, 256, 16, 4, ptr(x), ptr(x), ptr(x), ptr(x), ptr(x), ptr(x), ptr(x),
                                    torch.cuda.current_stream(dev).cuda_stream)
    assert rc != 0 and b"edge_mlp_fwd" in _lib.lib().gg_last_error()     # the sequencer falls back to two GEMM calls


# ------------------------------------------------------------------------------------------ layer
def _rand_graph(n, m, seed, isolated=0.1):
    from gnnome_assembly_b200.synth import make_random_graph
    return make_random_graph(n, m, seed=seed, isolated_frac=isolated)


@pytest.mark.parametrize("d", [64, 128, 256])
@pytest.mark.parametrize("batch_norm", [True, False])
def test_layer_matches_oracle(d, batch_norm, tc_mode):
    dev = _dev()
    O = _oracle()
    import gnnome_assembly_b200 as gg
    g = _rand_graph(300, 2500, seed=d)
    torch.manual_seed(d + int(batch_norm))
    ref = O.OracleGatedGCN(d, d, batch_norm)
    with torch.no_grad():
        for n_, p in ref.named_parameters():
            if "bn_" in n_:
                p.add_(0.3 * torch.randn_like(p))
    ours = gg.layers.GatedGCN_1d(d, d, batch_norm)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.to(dev)
    src = torch.from_numpy(g.src.astype(np.int64))
    dst = torch.from_numpy(g.dst.astype(np.int64))
    # pick inputs whose pre-ReLU values keep a margin from 0, so no ReLU mask can flip between the fp32
    # kernels and the fp64 oracle (a single flip moves a weight gradient by one edge's contribution)
    for attempt in range(50):
        h = torch.randn(g.num_nodes, d)
        e = torch.randn(g.num_edges, d)
        probe = {}
        with torch.no_grad():
            ref.double()(src, dst, g.num_nodes, h.double(), e.double(), probe=probe)
        if probe["min_abs_pre"] > 4e-6:
            break
    else:
        pytest.fail("could not find inputs with a ReLU margin")
    ref.float()
    gh = torch.randn(g.num_nodes, d)
    ge = torch.randn(g.num_edges, d)

    def run(mod, dev_, dtype, fn):
        mod = mod.to(dtype)
        hh = h.to(dev_, dtype).requires_grad_()
        ee = e.to(dev_, dtype).requires_grad_()
        ho, eo = fn(mod, hh, ee)
        (ho * gh.to(dev_, dtype)).sum().add((eo * ge.to(dev_, dtype)).sum()).backward()
        grads = {k: p.grad for k, p in mod.named_parameters()}
        grads["__h"], grads["__e"] = hh.grad, ee.grad
        return ho, eo, grads

    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    ho, eo, gr = run(ours, dev, torch.float32, lambda m_, hh, ee: m_(graph, hh, ee))
    ho64, eo64, gr64 = run(ref, "cpu", torch.float64, lambda m_, hh, ee: m_(src, dst, g.num_nodes, hh, ee))
    assert O.rel_err(ho, ho64) < 2e-5
    assert O.rel_err(eo, eo64) < 2e-5
    assert O.grads_close(gr, gr64, rtol=5e-4, atol_frac=2e-6) == []


@pytest.mark.parametrize("d,batch_norm", [(64, True), (128, True), (128, False), (256, True)])
def test_edge_gate_bulk_staged_kernel_is_bit_identical(d, batch_norm):
    """gg_debug_flags(8): the forward edge-gate pass with t / e_in staged through shared memory by cp.async.bulk
    (gg_layer_bulk.cuh) does the same arithmetic in the same per-node order as the default kernel."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200 import _lib
    g = _rand_graph(3000, 26000, seed=7 + d)            # several work blocks per CTA, chunks spanning nodes
    torch.manual_seed(d)
    layer = gg.layers.GatedGCN_1d(d, d, batch_norm).to(dev)
    graph = gg.AssemblyGraph(torch.from_numpy(g.src.astype(np.int64)), torch.from_numpy(g.dst.astype(np.int64)), g.num_nodes)
    h = torch.randn(g.num_nodes, d, device=dev)
    e = torch.randn(g.num_edges, d, device=dev)
    with torch.no_grad():
        h0, e0 = layer(graph, h, e)
        old = _lib.lib().gg_debug_flags(8)
        try:
            h1, e1 = layer(graph, h, e)
        finally:
            _lib.lib().gg_debug_flags(old)
    assert torch.equal(e0, e1) and torch.equal(h0, h1)


@pytest.mark.parametrize("d", [64, 128, 256])
def test_score_predictor_matches_oracle(d):
    dev = _dev()
    O = _oracle()
    import gnnome_assembly_b200 as gg
    g = _rand_graph(200, 1500, seed=7)
    torch.manual_seed(3)
    ref = O.OracleScorePredictor(d, 64)
    ours = gg.layers.ScorePredictor(d, 64)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.to(dev)
    src = torch.from_numpy(g.src.astype(np.int64))
    dst = torch.from_numpy(g.dst.astype(np.int64))
    x = torch.randn(g.num_nodes, d)
    e = torch.randn(g.num_edges, d)
    gs = torch.randn(g.num_edges, 1)
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    xd, ed = x.to(dev).requires_grad_(), e.to(dev).requires_grad_()
    s = ours(graph, xd, ed)
    assert s.shape == (g.num_edges, 1)
    (s * gs.to(dev)).sum().backward()
    ref = ref.double()
    x64, e64 = x.double().requires_grad_(), e.double().requires_grad_()
    s64 = ref(src, dst, x64, e64)
    (s64 * gs.double()).sum().backward()
    assert O.rel_err(s, s64) < 2e-5
    gr = {k: p.grad for k, p in ours.named_parameters()}
    gr64 = {k: p.grad for k, p in ref.named_parameters()}
    gr["__x"], gr["__e"], gr64["__x"], gr64["__e"] = xd.grad, ed.grad, x64.grad, e64.grad
    assert O.grads_close(gr, gr64, rtol=5e-4, atol_frac=2e-6) == []


# ------------------------------------------------------------------------------------------ model vs golden
def _run_model(g, state_dict, dev, grads=True):
    import gnnome_assembly_b200 as gg
    from oracle.gatedgcn_oracle import bce_loss
    model = gg.GraphGatedGCNModel(1, 2, g["d"], 16, g["L"], 64, g["batch_norm"], 16)
    model.load_state_dict(state_dict, strict=True)
    model.to(dev)
    graph = gg.AssemblyGraph(torch.from_numpy(g["src"].astype(np.int64)), torch.from_numpy(g["dst"].astype(np.int64)),
                             g["num_nodes"])
    e = torch.from_numpy(g["e"]).to(dev)
    pe = torch.from_numpy(g["pe"]).to(dev)
    x = torch.ones(g["num_nodes"], 1, device=dev)
    out = {}
    if grads:
        scores = model(graph, x, e, pe)
        loss = bce_loss(scores, torch.from_numpy(g["y"]).to(dev), g["pos_weight"])
        loss.backward()
        out["loss"] = float(loss.detach())
        out["grads"] = {k: p.grad for k, p in model.named_parameters()}
    else:
        with torch.no_grad():
            scores = model(graph, x, e, pe)
    out["scores"] = scores.detach()
    return out


@pytest.mark.parametrize("name", ["ref_rand_d64_L2_bn", "ref_rand_d64_L2_ln", "ref_asm_d128_L3_bn"])
def test_model_matches_reference_golden(golden_dir, name, tc_mode):
    dev = _dev()
    O = _oracle()
    g = torch.load(os.path.join(golden_dir, f"{name}.pt"), weights_only=False)
    out = _run_model(g, g["state_dict"], dev)
    assert out["scores"].shape == g["scores"].shape
    assert O.rel_err(out["scores"], g["scores"]) < TOL
    assert abs(out["loss"] - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
    bad = O.grads_close(out["grads"], g["grads"], rtol=2e-3, atol_frac=1e-5)
    if bad and tc_mode:
        # A pre-ReLU value that sits on the kink: ref_rand_d64_L2_bn has one edge-norm output of 1.1e-6 in layer 0
        # (tools/diag_golden_fp64.py prints the smallest ones from the fp64 oracle), the 3xTF32 products differ from
        # fp32 FFMA by a few 1e-7, so that ONE element's mask can flip and its gradient (5.7e-6) appears in / vanishes
        # from every sum it feeds.  Either side of a kink is a valid subgradient: accept a max-norm error of that
        # size when the tensor as a whole (Frobenius) still agrees to 5e-3.
        scale = max(float(v.abs().max()) for v in g["grads"].values())
        for k, err, _ in bad:
            r = g["grads"][k].double()
            fro = float((out["grads"][k].detach().double().cpu() - r).norm() / r.norm().clamp_min(1e-30))
            assert err <= 2e-4 * scale and fro <= 5e-3, (k, err, fro)
        bad = []
    assert bad == []


def test_model_checkpoint_golden(golden_dir, ckpt_path, tc_mode):
    """Config 3 of BASELINE.json at fixture size: shipped model_15xchr19.pt (d=256, L=16) on a chr21-like graph."""
    dev = _dev()
    O = _oracle()
    g = torch.load(os.path.join(golden_dir, "ref_asm_ckpt15xchr19.pt"), weights_only=False)
    sd = torch.load(ckpt_path, map_location="cpu")
    out = _run_model(g, sd, dev, grads=False)
    assert O.rel_err(out["scores"], g["scores"]) < TOL
    med = ((out["scores"].cpu() - g["scores"]).abs() / g["scores"].abs().clamp_min(1e-3)).median()
    assert float(med) < TOL


def test_model_eval_mode_and_no_grad_equal_train_mode(golden_dir):
    """BatchNorm1d(track_running_stats=False) uses batch statistics in eval too."""
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "ref_rand_d64_L2_bn.pt"), weights_only=False)
    a = _run_model(g, g["state_dict"], dev, grads=True)["scores"]
    b = _run_model(g, g["state_dict"], dev, grads=False)["scores"]
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------ full size
@pytest.fixture(scope="module")
def chr19_graph():
    from gnnome_assembly_b200.synth import make_assembly_graph
    return make_assembly_graph("chr19", seed=0)


def _full_model(dev, d=128, L=8, seed=0):
    import gnnome_assembly_b200 as gg
    torch.manual_seed(seed)
    m = gg.GraphGatedGCNModel(1, 2, d, 16, L, 64, True, 16)
    return m.to(dev)


def test_full_size_forward_vs_oracle(chr19_graph):
    """Config 2 workload (chr19-like graph, L=8, d=128): CUDA forward vs the fp32 CPU oracle."""
    dev = _dev()
    O = _oracle()
    import gnnome_assembly_b200 as gg
    g = chr19_graph
    model = _full_model(dev)
    oracle = O.OracleModel(1, 2, 128, 16, 8, 64, True, 16)
    oracle.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()}, strict=True)
    src = torch.from_numpy(g.src.astype(np.int64))
    dst = torch.from_numpy(g.dst.astype(np.int64))
    graph = gg.AssemblyGraph(src, dst, g.num_nodes)
    with torch.no_grad():
        s = model(graph, None, torch.from_numpy(g.e).to(dev), torch.from_numpy(g.pe).to(dev))
        ref = oracle(src, dst, g.num_nodes, torch.from_numpy(g.e), torch.from_numpy(g.pe))
    assert O.rel_err(s, ref) < TOL


def test_full_size_edge_order_invariance(chr19_graph):
    """Size-independent property: scores are attached to edge ids, so permuting the caller's edge order
    permutes the scores and nothing else (the plan's internal order must be invisible)."""
    dev = _dev()
    O = _oracle()
    import gnnome_assembly_b200 as gg
    g = chr19_graph
    model = _full_model(dev)
    perm = np.random.default_rng(1).permutation(g.num_edges)

    def run(src, dst, e):
        graph = gg.AssemblyGraph(torch.from_numpy(src.astype(np.int64)), torch.from_numpy(dst.astype(np.int64)),
                                 g.num_nodes)
        with torch.no_grad():
            return model(graph, None, torch.from_numpy(e).to(dev), torch.from_numpy(g.pe).to(dev))

    a = run(g.src, g.dst, g.e)
    b = run(g.src[perm], g.dst[perm], g.e[perm])
    assert O.rel_err(b, a[torch.from_numpy(perm).to(dev)]) < 2e-5
    c = run(g.src, g.dst, g.e)
    assert O.rel_err(c, a) < 1e-6                      # run-to-run (only fp64 stat atomics can reorder)


def test_full_size_training_step_finite_and_decreasing(chr19_graph):
    """train.py:245-258 on the CUDA path: a few Adam steps on one graph reduce the loss."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from oracle.gatedgcn_oracle import bce_loss
    g = chr19_graph
    model = _full_model(dev, d=64, L=2, seed=1)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    graph = gg.AssemblyGraph(torch.from_numpy(g.src.astype(np.int64)), torch.from_numpy(g.dst.astype(np.int64)),
                             g.num_nodes)
    e, pe, y = (torch.from_numpy(a).to(dev) for a in (g.e, g.pe, g.y))
    losses = []
    for _ in range(6):
        loss = bce_loss(model(graph, None, e, pe), y, 1 / 16.5)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_empty_and_isolated():
    """Edge cases: a graph with no edges at all, and nodes with zero in- or out-degree."""
    dev = _dev()
    O = _oracle()
    import gnnome_assembly_b200 as gg
    torch.manual_seed(0)
    model = gg.GraphGatedGCNModel(1, 2, 64, 16, 2, 64, True, 16).to(dev)
    oracle = O.OracleModel(1, 2, 64, 16, 2, 64, True, 16)
    oracle.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    # zero edges
    graph = gg.AssemblyGraph(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), 10)
    s = model(graph, None, torch.zeros(0, 2, device=dev), torch.randn(10, 18, device=dev))
    assert s.shape == (0, 1)
    # a star: node 0 has only out-edges, leaves have only in-edges, node 9 isolated
    src = torch.zeros(8, dtype=torch.int64)
    dst = torch.arange(1, 9)
    e, pe = torch.randn(8, 2), torch.randn(10, 18)
    s = model(gg.AssemblyGraph(src, dst, 10), None, e.to(dev), pe.to(dev))
    ref = oracle(src, dst, 10, e, pe)
    assert O.rel_err(s, ref) < TOL


def test_cpu_tensors_fail_loudly():
    _dev()
    import gnnome_assembly_b200 as gg
    model = gg.GraphGatedGCNModel(1, 2, 64, 16, 1, 64, True, 16).cuda()
    graph = gg.AssemblyGraph(torch.tensor([0, 1]), torch.tensor([1, 0]), 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(graph, None, torch.randn(2, 2), torch.randn(2, 18))
