"""Greedy contig decoding on the GPU — SURVEY.md §8f row 4; drop-in for inference.py:182-259 (`get_contigs`)
and inference.py:80-180 (`get_contigs_baselines`).

    walks = get_contigs(g, succs, preds, edges, nb_paths, len_threshold, device)      # inference.py:369 / :487

`g` needs what the reference reads: g.edata['score'], g.edata['prefix_length'], g.ndata['read_length'],
g.edges(), g.num_nodes().  succs / preds / edges are the reference's pickled dictionaries
(graph_parser.py:12-73); they may be None, in which case the adjacency is taken from g.edges() in edge-id
order — which is how graph_parser builds the dictionaries in the first place.

Per decoding iteration: sampling weights on the remaining graph (one kernel) -> nb_paths start edges (inverse
CDF on the device; `start_edges=` pins them for reproducible runs and parity tests) -> ALL walks of the
iteration concurrently, one warp each (`gg_decode_walks`) -> pick the walk reconstructing the longest sequence
-> `gg_decode_commit` adds it, its strand mates and the jumped-over nodes to the visited bitmap.  One small D2H
per iteration (lengths), one for the chosen walk.  There is no CPU path.

Deviations from the reference, on purpose:
  * self loops: the reference drops them with dgl.remove_self_loop (inference.py:187), which renumbers the edges while
    its `edges` dictionary keeps the old ids; here a self loop is left out of the successor / predecessor lists and gets
    sampling weight 0, and every other edge keeps its id;
  * a walk is cut, and the call raises, when it exceeds N nodes: the single-neighbour shortcut (inference.py:42-44,
    :65-67) skips the visited check, so on a cycle of single-neighbour nodes the reference never returns.  A finite walk
    that legitimately revisits nodes through that shortcut and grows past N nodes is reported the same way (the walk
    buffer holds 2N entries per walk: N forwards, N backwards);
  * when no edge is left to sample from, the loop ends; the reference raises (Categorical over zero edges).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda_device(device, *tensors):
    for t in tensors:
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    d = torch.device(device) if device is not None else None
    if d is not None and d.type == "cuda":
        return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())
    if not torch.cuda.is_available():
        raise RuntimeError("gnnome_assembly_b200.decode: no CUDA device (the decoder has no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())     # inference.py passes device='cpu' to get_contigs


def adjacency_arrays(src, dst, num_nodes):
    """(succ_ptr, succ_node, succ_eid, pred_ptr, pred_node, pred_eid, canon_eid) as int64 tensors on src's device.
    The reference fills its successor / predecessor lists by one pass over graph.edges() in edge-id order
    (graph_parser.py:27-29, :48-50) = a stable sort of the edge ids by src resp. dst.  It looks edges up through a
    {(src, dst): id} dictionary (graph_parser.py:69-72): of parallel edges only the LAST id is ever seen — scores and
    prefix lengths are read through that id (canon_eid)."""
    n, E = int(num_nodes), int(src.numel())
    ids = torch.arange(E, device=src.device, dtype=torch.int64)
    if E:
        _, inv = torch.unique(src * max(n, 1) + dst, return_inverse=True)
        last = torch.zeros(int(inv.max()) + 1, dtype=torch.int64, device=src.device).scatter_reduce_(0, inv, ids, "amax")
        canon = last[inv]
    else:
        canon = ids
    # self loops: the reference drops them before decoding (dgl.remove_self_loop, inference.py:187).  Here they are
    # left out of the successor / predecessor lists and get sampling weight 0 (gg_decode_edge_weights), while every
    # other edge KEEPS its id, so scores / prefix lengths stay addressed by the caller's edge ids.
    keep = src != dst
    ksrc, kdst, kcanon = src[keep], dst[keep], canon[keep]
    out = []
    for key, other in ((ksrc, kdst), (kdst, ksrc)):
        order = torch.argsort(key, stable=True)
        ptr_ = torch.zeros(n + 1, dtype=torch.int64, device=src.device)
        ptr_[1:] = torch.cumsum(torch.bincount(key, minlength=n), 0)
        out += [ptr_, other[order], kcanon[order]]
    return (*out, canon)


class DecodeGraph:
    """Successor / predecessor lists as device CSR in caller node ids, list order = the reference's."""

    def __init__(self, src, dst, num_nodes, device, succs=None, preds=None, edges=None):
        self.device = torch.device(device)
        self.num_nodes = int(num_nodes)
        src_t = torch.as_tensor(src).reshape(-1).to(torch.int64)
        dst_t = torch.as_tensor(dst).reshape(-1).to(torch.int64)
        self.num_edges = int(src_t.numel())
        self._src_t, self._dst_t = src_t, dst_t
        self._edge_of = None
        if succs is None or preds is None:
            arrays = adjacency_arrays(src_t.to(self.device), dst_t.to(self.device), self.num_nodes)
        else:
            src_np, dst_np = src_t.cpu().numpy(), dst_t.cpu().numpy()
            s_ptr, s_node, s_eid = self._from_dict(succs, edges, forward=True)
            p_ptr, p_node, p_eid = self._from_dict(preds, edges, forward=False)
            canon = np.array([edges[(int(a), int(b))] for a, b in zip(src_np, dst_np)], dtype=np.int64)
            arrays = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)) for a in
                      (s_ptr, s_node, s_eid, p_ptr, p_node, p_eid, canon)]
        i32 = lambda a: a.to(self.device, torch.int32).contiguous()
        (self.succ_ptr, self.succ_node, self.succ_eid, self.pred_ptr, self.pred_node, self.pred_eid,
         self.canon_eid) = (i32(a) for a in arrays)
        self.src, self.dst = i32(src_t), i32(dst_t)

    def _from_dict(self, adj, edges, forward):
        if edges is None:
            raise ValueError("decode: succs / preds given without the edges dictionary")
        ptr_, node, eid = [0], [], []
        for u in range(self.num_nodes):
            for v in adj.get(u, ()):
                node.append(v)
                eid.append(edges[(u, v)] if forward else edges[(v, u)])
            ptr_.append(len(node))
        return np.array(ptr_), np.array(node, dtype=np.int64), np.array(eid, dtype=np.int64)

    def edge_id(self, s, d):
        if self._edge_of is None:
            a, b = self._src_t.cpu().tolist(), self._dst_t.cpu().tolist()
            self._edge_of = {(u, v): i for i, (u, v) in enumerate(zip(a, b))}      # last id of a parallel pair wins
        return self._edge_of[(int(s), int(d))]


class WalkBatch:
    """Result of one decode_walks call (device tensors; `.host()` brings the small ones over in one go)."""

    def __init__(self, N, n_walks, device):
        words = (N + 31) // 32
        self.N, self.n, self.words = N, n_walks, words
        self.local_visited = torch.empty(max(n_walks, 1) * words, dtype=torch.int32, device=device)
        self.walk_buf = torch.empty(max(n_walks, 1) * 2 * N, dtype=torch.int32, device=device)
        self.meta = torch.zeros(2 * max(n_walks, 1) + 2, dtype=torch.int32, device=device)   # beg | len | err
        self.seq_len = torch.zeros(max(n_walks, 1), dtype=torch.int64, device=device)

    def host(self):
        meta = self.meta.cpu()
        n = self.n
        return meta[:n].numpy(), meta[n:2 * n].numpy(), self.seq_len.cpu().numpy()[:n], int(meta[2 * max(n, 1)])

    def walk(self, w, beg, length):
        return self.walk_buf[w * 2 * self.N + beg: w * 2 * self.N + beg + length]


def decode_walks(dg, scores, prefix_length, read_length, visited, start_src, start_dst, start_eid, out=None):
    """All walks of one iteration (inference.py:236-241 for every sampled edge).  Inputs are device tensors:
    scores fp32[E], prefix_length int64[E], read_length int64[N], visited int32[(N+31)//32] (bitmap),
    start_* int32[n_walks].  Returns a WalkBatch."""
    n = int(start_src.numel())
    wb = out if out is not None and out.n == n else WalkBatch(dg.num_nodes, n, dg.device)
    m = wb.meta
    check(_lib.lib().gg_decode_walks(
        dg.num_nodes, ptr(dg.succ_ptr), ptr(dg.succ_node), ptr(dg.succ_eid), ptr(dg.pred_ptr), ptr(dg.pred_node),
        ptr(dg.pred_eid), ptr(scores), ptr(prefix_length), ptr(read_length), ptr(visited), n,
        ptr(start_src), ptr(start_dst), ptr(start_eid), ptr(wb.local_visited), ptr(wb.walk_buf),
        m.data_ptr(), m.data_ptr() + 4 * n, ptr(wb.seq_len), m.data_ptr() + 4 * 2 * max(n, 1), _stream()),
        "gg_decode_walks")
    return wb


def commit_walk(dg, wb, w, beg, length, visited):
    """inference.py:223-234,241: the chosen walk, its strand mates and the jumped-over nodes become visited."""
    walk = wb.walk(w, beg, length)
    loc = wb.local_visited[w * wb.words:(w + 1) * wb.words]
    check(_lib.lib().gg_decode_commit(dg.num_nodes, ptr(dg.succ_ptr), ptr(dg.succ_node), ptr(dg.pred_ptr),
                                      ptr(dg.pred_node), ptr(walk), int(length), ptr(loc), ptr(visited), _stream()),
          "gg_decode_commit")


def sample_edges(dg, scores, visited, nb_paths, generator=None):
    """inference.py:279-286 on the graph without the visited nodes (:262-275): indices of nb_paths edges drawn
    with probability proportional to max(sigmoid(score), 1e-9).  None if no edge is left."""
    w = torch.empty(dg.num_edges, dtype=torch.float32, device=dg.device)
    check(_lib.lib().gg_decode_edge_weights(dg.num_edges, ptr(dg.src), ptr(dg.dst), ptr(scores), ptr(visited), ptr(w),
                                            _stream()), "gg_decode_edge_weights")
    cdf = torch.cumsum(w.double(), 0)
    total = float(cdf[-1]) if dg.num_edges else 0.0
    if total <= 0.0:
        return None
    u = torch.rand(nb_paths, dtype=torch.float64, device=dg.device, generator=generator) * total
    return torch.searchsorted(cdf, u, right=True).clamp_(max=dg.num_edges - 1)


def _decode_loop(g, succs, preds, edges, nb_paths, len_threshold, device, start_edges, score_list, generator):
    """The iteration of inference.py:182-259 / :80-180.  score_list[0] drives sampling, the choice of the best
    walk and the visited set; the other score arrays (the baselines of :134-141) are walked from the same start
    edges and reported at the chosen index."""
    dev = _cuda_device(device, *score_list)
    src, dst = g.edges()
    N = int(g.num_nodes())
    dg = DecodeGraph(src, dst, N, dev, succs, preds, edges)
    with torch.cuda.device(dev):
        scores_d = [torch.as_tensor(t).reshape(-1).to(dev, torch.float32).contiguous() for t in score_list]
        prefix = torch.as_tensor(g.edata["prefix_length"]).reshape(-1).to(dev, torch.int64).contiguous()
        rlen = torch.as_tensor(g.ndata["read_length"]).reshape(-1).to(dev, torch.int64).contiguous()
        visited = torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev)
        starts = iter(start_edges) if start_edges is not None else None
        results = [[] for _ in score_list]
        batches = [None] * len(score_list)
        while True:
            if starts is not None:
                try:
                    s_list, d_list = next(starts)
                except StopIteration:
                    break
                eid = torch.tensor([dg.edge_id(a, b) for a, b in zip(s_list, d_list)], dtype=torch.int32, device=dev)
                s_t = torch.as_tensor(np.asarray(s_list), dtype=torch.int32).to(dev)
                d_t = torch.as_tensor(np.asarray(d_list), dtype=torch.int32).to(dev)
            else:
                idx = sample_edges(dg, scores_d[0], visited, nb_paths, generator)
                if idx is None:
                    break
                eid = dg.canon_eid[idx]
                s_t, d_t = dg.src[idx], dg.dst[idx]
            for k, sc in enumerate(scores_d):
                batches[k] = decode_walks(dg, sc, prefix, rlen, visited, s_t, d_t, eid, out=batches[k])
            beg, length, seq_len, err = batches[0].host()
            if err:
                raise RuntimeError("decode: a greedy walk ran into a cycle of single-neighbour nodes "
                                   "(the reference's walk_forwards / walk_backwards would not terminate)")
            best = int(np.argmax(seq_len))                         # max(all_walks, key=get_contig_length): first maximum
            if int(length[best]) < len_threshold:                  # inference.py:243-244 / :164-165
                break
            results[0].append(batches[0].walk(best, int(beg[best]), int(length[best])).cpu().tolist())
            for k in range(1, len(scores_d)):
                bk, lk, _, errk = batches[k].host()
                if errk:
                    raise RuntimeError("decode: a baseline walk ran into a cycle of single-neighbour nodes")
                results[k].append(batches[k].walk(best, int(bk[best]), int(lk[best])).cpu().tolist())
            commit_walk(dg, batches[0], best, int(beg[best]), int(length[best]), visited)
    return results


def get_contigs(g, succs, preds, edges, nb_paths=50, len_threshold=20, device="cpu", start_edges=None, scores=None,
                generator=None):
    """Drop-in for inference.py:182-259.  Returns the list of contigs, each a list of node ids.
    start_edges: optional iterable of per-iteration (src_ids, dst_ids) sequences replacing the random draw.
    scores: optional per-edge tensor to walk on instead of g.edata['score']."""
    score_t = g.edata["score"] if scores is None else scores
    return _decode_loop(g, succs, preds, edges, nb_paths, len_threshold, device, start_edges, [score_t], generator)[0]


def get_contigs_baselines(g, succs, preds, edges, nb_paths=50, len_threshold=20, device="cpu", start_edges=None,
                          generator=None):
    """Drop-in for inference.py:80-180: besides the model's walks, the greedy walks on overlap length and on
    overlap similarity from the same start edges.  Returns (all_contigs, all_contigs_len, all_contigs_sim)."""
    res = _decode_loop(g, succs, preds, edges, nb_paths, len_threshold, device, start_edges,
                       [g.edata["score"], g.edata["overlap_length"], g.edata["overlap_similarity"]], generator)
    return res[0], res[1], res[2]
