"""Mini-batch (cluster) path, SURVEY.md §8f row 3: sub-graph plans built on the device by compaction of the
parent plan vs the numpy restatement (bit-exact: index work), the model on a sub-graph vs the CPU oracle on
the DGL-convention sub-graph, and the sampler / loader mirror of train.py:292-296."""
import numpy as np
import pytest
import torch

from oracle import subgraph_oracle as so


def _rand_graph(n, m, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n, m).astype(np.int64), rng.integers(0, n, m).astype(np.int64)


# ------------------------------------------------------------------------------------------ CPU: the oracle itself
@pytest.mark.parametrize("n,m,k", [(1, 0, 1), (6, 9, 3), (40, 300, 17), (200, 1500, 200)])
def test_oracle_node_subgraph_matches_loops(n, m, k):
    src, dst = _rand_graph(n, m, n + m)
    nodes = np.random.default_rng(k).permutation(n)[:k]
    a = so.node_subgraph(src, dst, n, nodes)
    b = so.node_subgraph_loops(src, dst, n, nodes)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_oracle_subplan_is_a_plan_of_the_subgraph():
    n, m = 300, 2500
    src, dst = _rand_graph(n, m, 5)
    node_perm = np.random.default_rng(1).permutation(n)
    parent = so.plan_arrays(src, dst, n, node_perm)
    nodes = np.random.default_rng(2).permutation(n)[:120]
    sub = so.subplan_arrays(parent, nodes)
    s, d, eid = so.node_subgraph(src, dst, n, nodes)
    assert np.array_equal(sub["parent_eid"], eid)
    # internal arrays describe the same edge set
    assert np.array_equal(sub["node_perm"][sub["src"]][sub["inv_perm"]], s)
    assert np.array_equal(sub["node_perm"][sub["dst"]][sub["inv_perm"]], d)
    assert np.all(np.diff(sub["dst"]) >= 0)
    # ... and the selected nodes keep the parent's internal order
    assert np.all(np.diff(parent["node_inv"][nodes[sub["node_perm"]]]) > 0)


def test_loader_batches_cpu():
    from gnnome_assembly_b200.minibatch import DataLoader
    dl = DataLoader(None, torch.arange(10), None, batch_size=4, shuffle=False, drop_last=False)
    assert [b.tolist() for b in dl.batches()] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]] and len(dl) == 3
    dl = DataLoader(None, torch.arange(10), None, batch_size=4, shuffle=True, drop_last=True)
    got = [b.tolist() for b in dl.batches()]
    assert len(got) == 2 == len(dl) and len({i for b in got for i in b}) == 8


# ------------------------------------------------------------------------------------------ GPU
def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _plan_np(plan):
    names = ["src", "dst", "in_ptr", "out_ptr", "out_eid", "perm", "inv_perm", "node_perm", "node_inv"]
    return {k: plan.array(k).cpu().numpy().astype(np.int64) for k in names}


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,k,relabel", [(1, 0, 1, True), (5, 7, 0, True), (96, 700, 40, False), (96, 700, 96, True),
                                           (5000, 40000, 1234, True), (60000, 500000, 25000, True)])
def test_subplan_arrays_bit_exact(n, m, k, relabel):
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan
    src, dst = _rand_graph(n, m, n + m + k)
    plan = GraphPlan(torch.from_numpy(src), torch.from_numpy(dst), n, dev, relabel=relabel)
    nodes = np.random.default_rng(k).permutation(n)[:k]
    sp = plan.subplan(torch.from_numpy(nodes))
    want = so.subplan_arrays(_plan_np(plan), nodes)
    assert sp.num_nodes == k and sp.num_edges == want["parent_eid"].size
    for name in ["src", "dst", "in_ptr", "out_ptr", "out_eid", "perm", "inv_perm", "node_perm", "node_inv",
                 "parent_eid", "csrc", "cdst"]:
        got = sp.array(name).cpu().numpy()
        assert np.array_equal(got, want[name]), name
    # a sub-plan is an ordinary plan: it can be cut again
    if k >= 2:
        nodes2 = np.random.default_rng(7).permutation(k)[:k // 2]
        sp2 = sp.subplan(torch.from_numpy(nodes2))
        want2 = so.subplan_arrays({**want}, nodes2)
        for name in ["src", "dst", "in_ptr", "out_ptr", "out_eid", "perm", "node_perm", "parent_eid"]:
            assert np.array_equal(sp2.array(name).cpu().numpy(), want2[name]), name


@pytest.mark.gpu
def test_subplan_rejects_bad_nodes():
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan
    src, dst = _rand_graph(50, 200, 0)
    plan = GraphPlan(torch.from_numpy(src), torch.from_numpy(dst), 50, dev)
    with pytest.raises(RuntimeError, match="out of range"):
        plan.subplan(torch.tensor([1, 2, 50]))
    with pytest.raises(RuntimeError, match="duplicate"):
        plan.subplan(torch.tensor([1, 2, 2]))
    assert plan.subplan(torch.tensor([3, 1])).num_nodes == 2        # the parent stays usable


@pytest.mark.gpu
def test_model_on_sampled_subgraph_matches_oracle():
    """train.py:292-306 on the engine: cluster sampler -> sub_g -> model(sub_g, x, e, pe) vs the CPU oracle on
    the DGL-convention sub-graph (<= 1e-4 relative on the logits, BASELINE.json's bar)."""
    dev = _dev()
    import gnnome_assembly_b200 as gg
    from gnnome_assembly_b200.minibatch import EID, NID, ClusterGCNSampler, DataLoader
    from gnnome_assembly_b200.synth import make_assembly_graph
    from oracle.gatedgcn_oracle import OracleModel, rel_err
    gs = make_assembly_graph("chr19", seed=3, genome_len=1_500_000)
    g = gg.AssemblyGraph(torch.from_numpy(gs.src.astype(np.int64)), torch.from_numpy(gs.dst.astype(np.int64)), gs.num_nodes)
    g.ndata["pe"] = torch.from_numpy(gs.pe)
    g.edata["e"] = torch.from_numpy(gs.e)
    g.edata["y"] = torch.from_numpy(gs.y)
    torch.manual_seed(0)
    oracle = OracleModel(1, 2, 64, 16, 2, 64, True, 16)
    model = gg.GraphGatedGCNModel(1, 2, 64, 16, 2, 64, True, 16)
    model.load_state_dict(oracle.state_dict())
    model.to(dev)
    k = 8
    sampler = ClusterGCNSampler(g, k, device=dev)
    loader = DataLoader(g, torch.arange(k), sampler, batch_size=3, shuffle=True, drop_last=False, num_workers=4)
    seen = 0
    for sub_g in loader:
        sub_g = sub_g.to(dev)
        nodes = sub_g.ndata[NID].cpu().numpy()
        s, d, eid = so.node_subgraph(gs.src, gs.dst, gs.num_nodes, nodes)
        assert np.array_equal(sub_g.edata[EID].cpu().numpy(), eid)
        assert np.array_equal(sub_g.edges()[0].cpu().numpy(), s) and np.array_equal(sub_g.edges()[1].cpu().numpy(), d)
        assert torch.equal(sub_g.edata["e"].cpu(), torch.from_numpy(gs.e[eid]))
        assert torch.equal(sub_g.ndata["pe"].cpu(), torch.from_numpy(gs.pe[nodes]))
        with torch.no_grad():
            got = model(sub_g, None, sub_g.edata["e"], sub_g.ndata["pe"]).cpu()
            ref = oracle(torch.from_numpy(s), torch.from_numpy(d), nodes.size, torch.from_numpy(gs.e[eid]),
                         torch.from_numpy(gs.pe[nodes]))
        assert rel_err(got, ref) < 1e-4
        seen += nodes.size
    assert seen == gs.num_nodes                                     # the clusters cover the graph exactly once
    # contiguous breadth-first chunks are a low-cut partition on an assembly graph
    part = sampler.partition_node_ids
    assert part.numel() == gs.num_nodes
