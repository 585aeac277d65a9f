"""CPU tests (no GPU): C-ABI exports, host-side logic, package surface, generator invariants, DP helper."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from gnnome_assembly_b200 import _lib
    header = open(os.path.join(ROOT, "include", "gnnome_b200.h")).read()
    declared = set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gnnome_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.gg_version() >= 100
    assert lib.gg_last_error() is not None


def test_sass_is_sm100a():
    import subprocess
    from gnnome_assembly_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    assert "sm_100a" in out.stdout


def test_sass_uses_tcgen05_tmem_and_tma():
    """The shipped library is Blackwell-native where it claims to be: the GEMM kernels issue tcgen05 MMAs
    (UTCHMMA) with TMEM loads / stores (LDTM / STTM), operands arrive through TMA tensor loads (UTMALDG), the fused
    g_t tile leaves through a TMA tensor store (UTMASTG); `profiles/r2_sass_summary.txt` is the per-kernel table."""
    import subprocess
    from gnnome_assembly_b200 import _lib
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "USETMAXREG"):
        assert mnemonic in sass, f"no {mnemonic} in the library's SASS"


def test_package_surface_matches_reference():
    import gnnome_assembly_b200 as gg
    for name in ("GatedGCN_1d", "GraphGatedGCN", "ScorePredictor", "NodeEncoder", "EdgeEncoder"):
        assert hasattr(gg.layers, name)
    assert hasattr(gg.models, "GraphGatedGCNModel")
    sys.path.insert(0, os.path.join(ROOT, "gnnome_assembly_b200", "dropin"))
    try:
        for m in ("layers", "models"):
            sys.modules.pop(m, None)
        import layers
        import models
        assert models.GraphGatedGCNModel is gg.models.GraphGatedGCNModel
        assert layers.GatedGCN_1d is gg.layers.GatedGCN_1d
    finally:
        sys.path.pop(0)
        for m in ("layers", "models"):
            sys.modules.pop(m, None)


def test_state_dict_keys_match_golden_and_checkpoint(golden_dir):
    import gnnome_assembly_b200 as gg
    g = torch.load(os.path.join(golden_dir, "ref_rand_d64_L2_bn.pt"), weights_only=False)
    m = gg.GraphGatedGCNModel(1, 2, 64, 16, 2, 64, True, 16)
    assert list(m.state_dict().keys()) == list(g["state_dict"].keys())
    m.load_state_dict(g["state_dict"], strict=True)
    ln = gg.GraphGatedGCNModel(1, 2, 64, 16, 2, 64, False, 16)
    g2 = torch.load(os.path.join(golden_dir, "ref_rand_d64_L2_ln.pt"), weights_only=False)
    ln.load_state_dict(g2["state_dict"], strict=True)


def test_checkpoint_loads_strict(ckpt_path):
    import gnnome_assembly_b200 as gg
    sd = torch.load(ckpt_path, map_location="cpu")
    m = gg.GraphGatedGCNModel(1, 2, 256, 16, 16, 64, True, 16)
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 6_390_961


def test_unsupported_sizes_raise():
    import gnnome_assembly_b200 as gg
    with pytest.raises(NotImplementedError):
        gg.layers.GatedGCN_1d(48, 48, True)
    with pytest.raises(NotImplementedError):
        gg.layers.ScorePredictor(128, 32)


def test_no_cpu_fallback():
    import gnnome_assembly_b200 as gg
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = gg.GraphGatedGCNModel(1, 2, 64, 16, 1, 64, True, 16)
    graph = gg.AssemblyGraph(torch.tensor([0, 1]), torch.tensor([1, 0]), 2)
    with pytest.raises(RuntimeError):
        m(graph, None, torch.randn(2, 2), torch.randn(2, 18))


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gnnome_assembly_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f


def test_synthetic_graph_invariants():
    from gnnome_assembly_b200.synth import make_assembly_graph
    g = make_assembly_graph("chr19", seed=3, genome_len=1_000_000)
    N, E = g.num_nodes, g.num_edges
    assert g.src.dtype == np.int32 and g.src.min() >= 0 and g.src.max() < N and g.dst.max() < N
    assert g.e.shape == (E, 2) and g.pe.shape == (N, 18) and g.y.shape == (E,)
    assert np.all(np.diff(g.src) >= 0)                                   # grouped by src
    assert abs(g.e[:, 0].mean()) < 1e-3 and abs(g.e[:, 0].std(ddof=1) - 1) < 1e-3
    # strand symmetry: (u -> v) present iff (v^1 -> u^1) present
    fwd = set(zip(g.src.tolist(), g.dst.tolist()))
    assert all(((v ^ 1), (u ^ 1)) in fwd for u, v in list(fwd)[:2000])
    assert np.array_equal(g.pe[:, 0], np.bincount(g.dst, minlength=N).astype(np.float32))
    assert np.array_equal(g.pe[:, 1], np.bincount(g.src, minlength=N).astype(np.float32))
    g2 = make_assembly_graph("chr19", seed=3, genome_len=1_000_000)
    assert np.array_equal(g.src, g2.src) and np.array_equal(g.e, g2.e)   # seeded


def test_chr19_scale():
    from gnnome_assembly_b200.synth import make_assembly_graph
    g = make_assembly_graph("chr19", seed=0)
    assert 40_000 < g.num_nodes < 52_000 and 330_000 < g.num_edges < 400_000


def test_shard_indices():
    from gnnome_assembly_b200.dp import shard_indices
    got = [shard_indices(15, r, 8) for r in range(8)]
    assert got[0] == [0, 8] and got[6] == [6, 14] and got[7] == [7, None]
    assert sorted(i for w in got for i in w if i is not None) == list(range(15))


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from gnnome_assembly_b200.dp import GradBucket
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    bucket = GradBucket(lin.parameters())
    bucket.allreduce_mean(active=True)
    a = lin.weight.grad.clone()
    # short wave: rank 1 idle -> mean over the single active rank
    lin.zero_grad()
    lin(x).sum().backward()
    bucket.allreduce_mean(active=(rank == 0))
    q.put((rank, a, lin.weight.grad.clone()))
    dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    # rank r's weight grad rows are 2*(r+1) each -> mean = 3 ; short wave -> rank 0's own grad = 2
    for _, a, b in res:
        assert torch.allclose(a, torch.full((3, 4), 3.0))
        assert torch.allclose(b, torch.full((3, 4), 2.0))


def _arena_worker(rank, world, port, q):
    import torch.distributed as dist
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200.dp import ArenaSync
    from gnnome_assembly_b200.flat import GradArena, ensure_flat
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = gg.GraphGatedGCNModel(1, 2, 64, 16, 2, 64, True, 16)          # host-resident: only the sync logic runs
    sync = ArenaSync(model, bucket_layers=1)          # conv1 | conv0 + head: two collectives per step
    layout = ensure_flat(model)
    assert [b[-1] for b in sync._schedule(layout)] == [["conv1"], ["conv0", "head"]]

    def fake_backward(value):
        """what a backward pass does to the arena: fill, hand autograd-style views to .grad, signal the layers"""
        arena = GradArena(layout, torch.device("cpu"))
        model.arena_hook(arena)
        arena.tensor().fill_(value)
        for p, off in layout.entries:
            p.grad = arena.tensor()[off:off + p.numel()].view_as(p)
        arena.segment_done("conv1")                                         # backward order: last layer first
        arena.segment_done("conv0")
        return arena

    sync.begin(world)
    fake_backward(float(rank + 1))
    sync.finish()
    a = [float(p.grad.flatten()[0]) for p in model.parameters()]
    # short wave: rank 1 idle, divisor = 1 active rank
    sync.begin(1)
    if rank == 0:
        fake_backward(5.0)
        sync.finish()
    else:
        sync.idle_step()
    b = [float(p.grad.flatten()[0]) for p in model.parameters()]
    # autograd made private copies instead of adopting the arena views (seen under compute-sanitizer): the copies hold
    # the local values, the arena the mean -> finish() re-points .grad into the arena
    sync.begin(world)
    for p in model.parameters():
        p.grad = None
    arena = GradArena(layout, torch.device("cpu"))
    model.arena_hook(arena)
    arena.tensor().fill_(float(10 * (rank + 1)))
    for p, off in layout.entries:
        p.grad = arena.tensor()[off:off + p.numel()].view_as(p).clone()
    arena.segment_done("conv1")
    arena.segment_done("conv0")
    sync.finish()
    base = arena.tensor().data_ptr()
    assert all(p.grad.data_ptr() == base + 4 * off for p, off in layout.entries)
    c = [float(p.grad.flatten()[0]) for p in model.parameters()]
    # ... but a .grad that existed before the backward (accumulation) cannot be replaced: loud error, on every rank
    sync.begin(world)
    arena = GradArena(layout, torch.device("cpu"))
    model.arena_hook(arena)
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    arena.tensor().fill_(1.0)
    arena.segment_done("conv1")
    arena.segment_done("conv0")
    try:
        sync.finish()
        refused = False
    except RuntimeError as err:
        refused = "zero_grad" in str(err)
    q.put((rank, a, b, c, refused))
    dist.destroy_process_group()


def test_arena_sync_gloo_world2():
    """dp.ArenaSync: per-segment all-reduce of the flat gradient arena (layer segments as their backward finishes, the
    head segment in finish), mean over the active ranks, an idle rank of a short wave walks the same collectives."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_arena_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for _, a, b, c, refused in res:
        assert all(abs(x - 1.5) < 1e-6 for x in a) and len(a) == 42          # mean of 1 and 2, every parameter
        assert all(abs(x - 5.0) < 1e-6 for x in b)                           # rank 0's gradient alone, on both ranks
        assert all(abs(x - 15.0) < 1e-6 for x in c)                          # copied gradients re-pointed to the mean
        assert refused


def test_flat_parameters_keep_the_reference_state_dict():
    """flat.ensure_flat: parameters become views of one buffer (Wn / bn are strided views, no cat), values, names and
    shapes of the state_dict are untouched, deep copies and load_state_dict keep working."""
    import copy
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200.flat import GradArena, ensure_flat, packed_node_weights
    torch.manual_seed(1)
    m = gg.GraphGatedGCNModel(1, 2, 64, 16, 2, 64, True, 16)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    layout = ensure_flat(m)
    assert [n for n, _, _ in layout.segments] == ["head", "conv0", "conv1"]
    assert list(m.state_dict().keys()) == list(sd0.keys())
    assert all(torch.equal(sd0[k], v) for k, v in m.state_dict().items())
    c = m.gnn.convs[1]
    Wn, bn = packed_node_weights(c)
    assert torch.equal(Wn, torch.cat([c.A_1.weight, c.A_2.weight, c.A_3.weight, c.B_1.weight, c.B_2.weight], 0))
    assert torch.equal(bn, torch.cat([c.A_1.bias, c.A_2.bias, c.A_3.bias, c.B_1.bias, c.B_2.bias], 0))
    assert Wn.data_ptr() == c.A_1.weight.data_ptr()                          # a view, not a copy
    assert ensure_flat(m) is layout                                          # idempotent
    with torch.no_grad():
        c.A_3.weight.add_(1.0)                                               # an optimizer update through the parameter
    assert torch.equal(packed_node_weights(c)[0][128:192], c.A_3.weight)     # ... is seen by the packed view
    m2 = copy.deepcopy(m)                                                    # train.py:199-206 keeps a best_model copy
    assert ensure_flat(m2) is not layout and packed_node_weights(m2.gnn.convs[0]) is not None
    m.load_state_dict(sd0, strict=True)
    assert ensure_flat(m) is layout and torch.equal(c.A_3.weight, sd0["gnn.convs.1.A_3.weight"])
    arena = GradArena(layout, torch.device("cpu"))
    assert arena.slot(c.A_1.weight, rows=5 * 64 * 64).numel() == 5 * 64 * 64
    assert arena.slot(c.bn_h.bias).data_ptr() == arena.tensor().data_ptr() + 4 * layout.offset_of[id(c.bn_h.bias)]


# ------------------------------------------------------------------------------------------ graph files, dgl stand-in
def test_graph_file_round_trip(tmp_path):
    """DGL-free graph container (graph_io.py) behind dgl.save_graphs / dgl.load_graphs (graph_dataset.py:72,129)."""
    from gnnome_assembly_b200.graph import AssemblyGraph
    from gnnome_assembly_b200.graph_io import load_graphs, save_graphs
    rng = np.random.default_rng(0)
    graphs = []
    for n, m, idt in ((5, 0, torch.int64), (40, 300, torch.int32), (1000, 7000, torch.int64)):
        g = AssemblyGraph(torch.from_numpy(rng.integers(0, n, m)).to(idt), torch.from_numpy(rng.integers(0, n, m)).to(idt), n)
        g.ndata["read_length"] = torch.from_numpy(rng.integers(1000, 30000, n))
        g.ndata["pe"] = torch.randn(n, 3)
        g.edata["overlap_similarity"] = torch.rand(m)
        g.edata["y"] = (torch.rand(m) > 0.5).float()
        graphs.append(g)
    p = str(tmp_path / "7.dgl")
    save_graphs(p, graphs, {"glabel": torch.arange(3)})
    back, labels = load_graphs(p)
    assert torch.equal(labels["glabel"], torch.arange(3)) and len(back) == 3
    for a, b in zip(graphs, back):
        assert a.num_nodes() == b.num_nodes() and a.edges()[0].dtype == b.edges()[0].dtype
        assert torch.equal(a.edges()[0], b.edges()[0]) and torch.equal(a.edges()[1], b.edges()[1])
        for k in a.ndata:
            assert torch.equal(a.ndata[k], b.ndata[k]) and a.ndata[k].dtype == b.ndata[k].dtype
        for k in a.edata:
            assert torch.equal(a.edata[k], b.edata[k])
    (only,), _ = load_graphs(p, idx_list=[1])
    assert only.num_nodes() == 40
    with open(str(tmp_path / "bad.dgl"), "wb") as f:
        f.write(b"\x00" * 64)
    with pytest.raises(ValueError, match="not a gnnome_assembly_b200 graph file"):
        load_graphs(str(tmp_path / "bad.dgl"))


def test_dgl_standin_surface():
    """dropin/dgl covers the names the reference's loops touch (train.py:17-18,172,292-293; utils.py:34,68,102-124;
    graph_dataset.py:5,72,129; inference.py:184,271-273) and follows DGL's sub-graph conventions."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "dgl_standin", os.path.join(ROOT, "gnnome_assembly_b200", "dropin", "dgl", "__init__.py"),
        submodule_search_locations=[os.path.join(ROOT, "gnnome_assembly_b200", "dropin", "dgl")])
    dgl = importlib.util.module_from_spec(spec)
    sys.modules["dgl_standin"] = dgl
    spec.loader.exec_module(dgl)
    for name in ("graph", "seed", "load_graphs", "save_graphs", "remove_self_loop", "node_subgraph", "reverse", "NID", "EID"):
        assert hasattr(dgl, name), name
    for name in ("ClusterGCNSampler", "DataLoader", "MultiLayerFullNeighborSampler", "GraphDataLoader"):
        assert hasattr(dgl.dataloading, name), name
    assert hasattr(dgl.data, "DGLDataset")
    g = dgl.graph((torch.tensor([0, 1, 2, 2, 3]), torch.tensor([1, 2, 0, 2, 1])), num_nodes=5)
    g.edata["w"] = torch.arange(5.0)
    g.ndata["h"] = torch.arange(5.0)
    assert g.int().edges()[0].dtype == torch.int32 and g.int().long().edges()[0].dtype == torch.int64
    assert g.in_degrees().tolist() == [1, 2, 2, 0, 0] and g.out_degrees().tolist() == [1, 1, 2, 1, 0]
    A = g.adjacency_matrix(scipy_fmt="csr")
    assert A.shape == (5, 5) and A[2, 2] == 1 and A[3, 1] == 1 and A.sum() == 5
    r = dgl.remove_self_loop(g)
    assert r.num_edges() == 4 and r.edata["w"].tolist() == [0.0, 1.0, 2.0, 4.0]
    s = dgl.node_subgraph(g, torch.tensor([2, 0, 1]))
    assert s.ndata[dgl.NID].tolist() == [2, 0, 1] and s.edata[dgl.EID].tolist() == [0, 1, 2, 3]
    assert s.edges()[0].tolist() == [1, 2, 0, 0] and s.edges()[1].tolist() == [2, 0, 1, 0]
    assert s.ndata["h"].tolist() == [2.0, 0.0, 1.0]
    rv = dgl.reverse(g, copy_edata=True)
    assert rv.edges()[0].tolist() == g.edges()[1].tolist() and rv.edata["w"].tolist() == g.edata["w"].tolist()
    # moving / casting keeps the structure: the derived graph points at its origin, whose GraphPlan it will share
    assert g.int()._origin is g and g.int().long()._origin is g


def test_plan_fingerprint_is_structural():
    """plan.py: the second-level plan cache is keyed by the edge list's content, so `g.to(device)` returning a new
    object every step (train.py:243) does not rebuild the plan."""
    from gnnome_assembly_b200.plan import _fingerprint
    rng = np.random.default_rng(1)
    s, d = torch.from_numpy(rng.integers(0, 50, 400)), torch.from_numpy(rng.integers(0, 50, 400))
    k = _fingerprint(s, d, 50, "cuda:0")
    assert k == _fingerprint(s.clone().int(), d.clone().int(), 50, "cuda:0")
    perm = torch.from_numpy(rng.permutation(400))
    assert k != _fingerprint(s[perm], d[perm], 50, "cuda:0")           # the internal order depends on the edge order
    assert k != _fingerprint(d, s, 50, "cuda:0") and k != _fingerprint(s, d, 51, "cuda:0")
    assert k != _fingerprint(s, d, 50, "cuda:1")
    s2 = s.clone()
    s2[17] = (s2[17] + 1) % 50
    assert k != _fingerprint(s2, d, 50, "cuda:0")
