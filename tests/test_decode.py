"""Greedy contig decoding, SURVEY.md §8f row 4.  The golden vectors (tests/golden/ref_decode_*.pt) were
produced by the reference's own inference.get_contigs (make_golden_decode.py); the oracle is pinned to them on
the CPU, the CUDA decoder is held to them and to the oracle bit for bit (integer / comparison work)."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as do

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


# ------------------------------------------------------------------------------------------ CPU: pin the oracle
@pytest.mark.parametrize("name", ["ref_decode_small", "ref_decode_asm"])
def test_oracle_reproduces_reference_contigs(name):
    g = _gold(name)
    succs, preds, edge_id = do.adjacency(g["src"], g["dst"], g["num_nodes"])
    contigs, _ = do.get_contigs(g["score"], g["prefix_length"], g["read_length"], succs, preds, edge_id,
                                g["start_edges"], g["len_threshold"])
    assert contigs == g["contigs"]


@pytest.mark.parametrize("name", ["ref_decode_small", "ref_decode_asm"])
def test_oracle_reproduces_reference_baselines(name):
    g = _gold(name)
    b = g["baselines"]
    succs, preds, edge_id = do.adjacency(g["src"], g["dst"], g["num_nodes"])
    got = do.get_contigs_baselines([g["score"], g["overlap_length"], g["overlap_similarity"]], g["prefix_length"],
                                   g["read_length"], succs, preds, edge_id, b["start_edges"], g["len_threshold"])
    assert got[0] == b["contigs"] and got[1] == b["contigs_len"] and got[2] == b["contigs_sim"]


@pytest.mark.parametrize("n,m", [(1, 0), (6, 9), (50, 400), (300, 5000)])
def test_adjacency_arrays_match_reference_dictionaries(n, m):
    """host logic of the decoder (device-agnostic torch code, run on the CPU here): CSR lists in the order of the
    reference's dictionaries, parallel edges resolved to the last id like the reference's {(src,dst): id} lookup"""
    from gnnome_assembly_b200.decode import adjacency_arrays
    rng = np.random.default_rng(n + m)
    src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    sp, sn, se, pp, pn, pe, canon = [a.numpy() for a in adjacency_arrays(torch.from_numpy(src), torch.from_numpy(dst), n)]
    succs, preds, eid = do.adjacency(src, dst, n)
    for u in range(n):
        assert sn[sp[u]:sp[u + 1]].tolist() == succs[u] and pn[pp[u]:pp[u + 1]].tolist() == preds[u]
        assert se[sp[u]:sp[u + 1]].tolist() == [eid[(u, v)] for v in succs[u]]
        assert pe[pp[u]:pp[u + 1]].tolist() == [eid[(v, u)] for v in preds[u]]
    assert canon.tolist() == [eid[(a, b)] for a, b in zip(src.tolist(), dst.tolist())]
    # self loops are dropped from the lists (dgl.remove_self_loop, inference.py:187) but every other edge keeps its id
    sp, sn, se, pp, pn, pe, canon = [a.numpy() for a in adjacency_arrays(torch.tensor([0, 1, 1]), torch.tensor([1, 1, 0]), 2)]
    assert sp.tolist() == [0, 1, 2] and sn.tolist() == [1, 0] and se.tolist() == [0, 2]
    assert pp.tolist() == [0, 1, 2] and pn.tolist() == [1, 0] and pe.tolist() == [2, 0]
    assert canon.tolist() == [0, 1, 2]


def test_oracle_walk_semantics_small():
    # 0 -> 2 -> 4 -> 6 with a tempting branch 2 -> 8 (higher score) that is already visited
    src = np.array([0, 2, 2, 4]); dst = np.array([2, 4, 8, 6])
    succs, preds, eid = do.adjacency(src, dst, 10)
    scores = np.array([0.0, 0.1, 0.9, 0.0], dtype=np.float32)
    w, seen = do.greedy_walk(0, scores, succs, eid, set(), True)
    assert w == [0, 2, 8] and seen == {0, 1, 2, 3, 8, 9}
    w, _ = do.greedy_walk(0, scores, succs, eid, {8}, True)
    assert w == [0, 2, 4, 6]
    w, _ = do.greedy_walk(6, scores, preds, eid, set(), False)
    assert w == [0, 2, 4, 6]
    assert do.contig_length([0, 2, 4, 6], np.array([5, 7, 100, 11]), np.arange(10) * 1000, eid) == 5 + 7 + 11 + 6000


# ------------------------------------------------------------------------------------------ GPU
def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _graph(gold_or_synth, score):
    import gnnome_assembly_b200 as gg
    s = gold_or_synth
    get = (lambda k: s[k]) if isinstance(s, dict) else (lambda k: getattr(s, k))
    g = gg.AssemblyGraph(torch.from_numpy(np.asarray(get("src"), dtype=np.int64)),
                         torch.from_numpy(np.asarray(get("dst"), dtype=np.int64)), int(get("num_nodes")))
    g.edata["score"] = torch.from_numpy(np.asarray(score, dtype=np.float32))
    g.edata["prefix_length"] = torch.from_numpy(np.asarray(get("prefix_length"), dtype=np.int64))
    g.ndata["read_length"] = torch.from_numpy(np.asarray(get("read_length"), dtype=np.int64))
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_decode_small", "ref_decode_asm"])
@pytest.mark.parametrize("with_dicts", [False, True])
def test_gpu_get_contigs_matches_reference_golden(name, with_dicts):
    _dev()
    from gnnome_assembly_b200.decode import get_contigs
    gold = _gold(name)
    g = _graph(gold, gold["score"])
    succs = preds = edges = None
    if with_dicts:
        succs, preds, edges = do.adjacency(gold["src"], gold["dst"], gold["num_nodes"])
    contigs = get_contigs(g, succs, preds, edges, gold["nb_paths"], gold["len_threshold"], device="cpu",
                          start_edges=gold["start_edges"])
    assert contigs == gold["contigs"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_decode_small", "ref_decode_asm"])
def test_gpu_get_contigs_baselines_matches_reference_golden(name):
    _dev()
    from gnnome_assembly_b200.decode import get_contigs_baselines
    gold = _gold(name)
    b = gold["baselines"]
    g = _graph(gold, gold["score"])
    g.edata["overlap_length"] = torch.from_numpy(np.asarray(gold["overlap_length"], dtype=np.float32))
    g.edata["overlap_similarity"] = torch.from_numpy(np.asarray(gold["overlap_similarity"], dtype=np.float32))
    got = get_contigs_baselines(g, None, None, None, gold["nb_paths"], gold["len_threshold"], device="cpu",
                                start_edges=b["start_edges"])
    assert got[0] == b["contigs"] and got[1] == b["contigs_len"] and got[2] == b["contigs_sim"]


@pytest.mark.gpu
def test_gpu_walks_match_oracle_with_visited_nodes():
    dev = _dev()
    from gnnome_assembly_b200 import decode
    from gnnome_assembly_b200.synth import make_assembly_graph
    gs = make_assembly_graph("chr19", seed=5, genome_len=6_000_000, pe_dim=0)
    rng = np.random.default_rng(0)
    score = ((gs.y * 2 - 1) * 2 + rng.standard_normal(gs.num_edges) * 2).astype(np.float32)
    succs, preds, eid = do.adjacency(gs.src, gs.dst, gs.num_nodes)
    vis_nodes = rng.choice(gs.num_nodes // 2, gs.num_nodes // 10, replace=False) * 2
    visited = set(vis_nodes.tolist()) | set((vis_nodes + 1).tolist())
    free = np.nonzero(~np.isin(gs.src, list(visited)) & ~np.isin(gs.dst, list(visited)))[0]
    pick = rng.choice(free, 64, replace=True)
    starts = list(zip(gs.src[pick].tolist(), gs.dst[pick].tolist()))
    walks, visiteds = do.walks_for_starts(starts, score, succs, preds, eid, visited)
    lengths = [do.contig_length(w, gs.prefix_length, gs.read_length, eid) for w in walks]

    dg = decode.DecodeGraph(gs.src, gs.dst, gs.num_nodes, dev)
    bm = np.zeros((gs.num_nodes + 31) // 32, dtype=np.uint32)
    for v in visited:
        bm[v >> 5] |= np.uint32(1 << (v & 31))
    vis_t = torch.from_numpy(bm.view(np.int32)).to(dev)
    i32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int32)).to(dev)
    wb = decode.decode_walks(dg, torch.from_numpy(score).to(dev), torch.from_numpy(gs.prefix_length).to(dev),
                             torch.from_numpy(gs.read_length).to(dev), vis_t, i32(gs.src[pick]), i32(gs.dst[pick]),
                             dg.canon_eid[torch.from_numpy(pick).to(dev)])
    beg, length, seq_len, err = wb.host()
    assert err == 0
    loc = wb.local_visited.cpu().numpy().view(np.uint32).reshape(-1, wb.words)
    for w in range(len(starts)):
        assert wb.walk(w, int(beg[w]), int(length[w])).cpu().tolist() == walks[w], w
        assert int(seq_len[w]) == lengths[w]
        got = {32 * i + b for i in np.nonzero(loc[w])[0] for b in range(32) if (loc[w][i] >> np.uint32(b)) & 1}
        assert got == visiteds[w]
    # sampling weights on the remaining graph (before the commit)
    from gnnome_assembly_b200 import _lib
    score_d = torch.from_numpy(score).to(dev)
    wts = torch.empty(gs.num_edges, device=dev)
    _lib.check(_lib.lib().gg_decode_edge_weights(gs.num_edges, dg.src.data_ptr(), dg.dst.data_ptr(), score_d.data_ptr(),
                                                 vis_t.data_ptr(), wts.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "gg_decode_edge_weights")
    ref_w = do.edge_weights(gs.src, gs.dst, score, visited)
    assert 0 < (ref_w > 0).sum() < gs.num_edges
    assert np.array_equal(wts.cpu().numpy() == 0, ref_w == 0)
    assert np.allclose(wts.cpu().numpy(), ref_w, rtol=2e-6, atol=0)
    idx = decode.sample_edges(dg, score_d, vis_t, 500)
    assert np.all(ref_w[idx.cpu().numpy()] > 0)
    # commit: walk + mates + jumped-over nodes
    best = int(np.argmax(lengths))
    decode.commit_walk(dg, wb, best, int(beg[best]), int(length[best]), vis_t)
    want = set(visited) | visiteds[best]
    for a, b in zip(walks[best][:-1], walks[best][1:]):
        t = set(succs[a]) & set(preds[b])
        want |= t | {x ^ 1 for x in t}
    after = vis_t.cpu().numpy().view(np.uint32)
    got = {32 * i + b for i in np.nonzero(after)[0] for b in range(32) if (after[i] >> np.uint32(b)) & 1}
    assert got == want
    # after the commit nothing is left to sample from on this single-chromosome graph
    assert decode.sample_edges(dg, score_d, vis_t, 10) is None or want != set(range(gs.num_nodes))


@pytest.mark.gpu
def test_gpu_decode_full_size_properties():
    """chr19-size graph, random sampler: every contig is a path of the graph, no node (or strand mate) is used
    twice across contigs, and the whole decode takes seconds."""
    _dev()
    import time
    from gnnome_assembly_b200.decode import get_contigs
    from gnnome_assembly_b200.synth import make_assembly_graph
    gs = make_assembly_graph("chr19", seed=0, pe_dim=0)
    rng = np.random.default_rng(1)
    score = ((gs.y * 2 - 1) * 3 + rng.standard_normal(gs.num_edges)).astype(np.float32)
    g = _graph(gs, score)
    gen = torch.Generator(device="cuda").manual_seed(0)
    t0 = time.perf_counter()
    contigs = get_contigs(g, None, None, None, nb_paths=50, len_threshold=20, device="cuda", generator=gen)
    dt = time.perf_counter() - t0
    assert len(contigs) >= 1 and dt < 120
    edges = set(zip(gs.src.tolist(), gs.dst.tolist()))
    used = set()
    for c in contigs:
        assert len(c) >= 20
        assert all((a, b) in edges for a, b in zip(c[:-1], c[1:]))
        reads = {n >> 1 for n in c}
        assert len(reads) == len(c) and not (reads & used)
        used |= reads
    assert sum(len(c) for c in contigs) > 0.2 * gs.num_nodes / 2 * 0.5


@pytest.mark.gpu
def test_gpu_decode_drops_self_loops():
    """A self loop never enters a walk and is never sampled (inference.py:187); the other edges keep their ids."""
    _dev()
    from gnnome_assembly_b200 import decode
    dg = decode.DecodeGraph(np.array([0, 1, 1]), np.array([1, 1, 0]), 2, "cuda:0")
    assert dg.succ_node.tolist() == [1, 0] and dg.succ_eid.tolist() == [0, 2]
    vis = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    w = torch.empty(3, device="cuda:0")
    sc = torch.zeros(3, device="cuda:0")
    from gnnome_assembly_b200 import _lib
    _lib.check(_lib.lib().gg_decode_edge_weights(3, dg.src.data_ptr(), dg.dst.data_ptr(), sc.data_ptr(), vis.data_ptr(),
                                                 w.data_ptr(), torch.cuda.current_stream().cuda_stream), "w")
    assert w.tolist() == [0.5, 0.0, 0.5]
