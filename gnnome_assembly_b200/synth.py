"""Synthetic assembly (overlap) graph generator — the workload of every config in BASELINE.json.

The reference ships no graph data (real graphs = 43 GB download, download_dataset.sh); its graphs
come from simulated HiFi reads (pipeline.py:142-170) assembled by Raven (graph_dataset.py:96-126).
This module restates the idealised, error-free version of that process: reads with
the reference's own length distribution at 32.4x coverage, contained reads dropped, a + strand edge
i->j iff start_i < start_j < end_i - 500, every edge mirrored on the reverse-complement strand
(+ node 2k, - node 2k+1, as in inference.py:39 / algorithms.py:139), a few false "repeat" edges.
Edge features / positional encoding follow utils.py:67-75 and utils.py:97-138.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

CHR_LEN = {"chr19": 61_707_364, "chr21": 45_090_682}       # pipeline.py:38,40
COVERAGE = 32.4
_QPATH = os.path.join(os.path.dirname(__file__), "data", "read_length_quantiles.npz")


@dataclass
class SynthGraph:
    src: np.ndarray          # int32 [E], edge-id order (grouped by src, graph_parser.py:297)
    dst: np.ndarray          # int32 [E]
    num_nodes: int
    e: np.ndarray            # float32 [E,2]  z-scored overlap_length, overlap_similarity (utils.py:70-74)
    pe: np.ndarray           # float32 [N,18] in_deg | out_deg | 16-step PageRank (train.py:249-251)
    y: np.ndarray            # float32 [E]    1 = true overlap
    overlap_length: np.ndarray
    overlap_similarity: np.ndarray
    read_length: np.ndarray = None      # int64 [N]  ndata['read_length']  (graph_parser.py:256,284)
    prefix_length: np.ndarray = None    # int64 [E]  edata['prefix_length'] = read_length[src] - overlap_length (:280-281)

    @property
    def num_edges(self):
        return int(self.src.shape[0])


def _sample_lengths(rng, chrom, n):
    q = np.load(_QPATH)[chrom]
    u = rng.random(n) * (len(q) - 1)
    i = np.minimum(u.astype(np.int64), len(q) - 2)
    return np.maximum((q[i] + (u - i) * (q[i + 1] - q[i])).astype(np.int64), 1000)


def pagerank_pe(src, dst, n, pe_dim=16, alpha=0.95):
    """utils.py:124-138 (type_pe == 'PR'): x <- alpha * (D^-1 A)^T x + (1-alpha)/n, pe_dim steps."""
    import scipy.sparse as sp
    A = sp.csr_matrix((np.ones(len(src)), (src, dst)), shape=(n, n))
    D = np.asarray(A.sum(axis=1)).squeeze()
    Dinv = 1.0 / (D + 1e-9)
    Dinv[D < 1e-9] = 0
    P = (sp.diags(Dinv) @ A).T.tocsr()
    x = np.ones(n) / n
    cols = []
    for _ in range(pe_dim):
        x = alpha * P.dot(x) + (1.0 - alpha) / n
        cols.append(x.astype(np.float32))
    return np.stack(cols, axis=-1)


def make_assembly_graph(chrom="chr19", seed=0, genome_len=None, target_edges=None, p_fp=0.05,
                        pe_dim=16, min_overlap=500) -> SynthGraph:
    """Generate a chr-like assembly graph.  `target_edges` rescales the genome length so that
    E ~= target_edges (config 5 of BASELINE.json: 1M / 5M / 20M edges)."""
    rng = np.random.default_rng(seed)
    G = int(genome_len if genome_len is not None else CHR_LEN[chrom])
    if target_edges is not None:
        G = int(CHR_LEN["chr19"] * (target_edges / 355_768.0))
    mean_len = float(np.load(_QPATH)[chrom].mean())
    n_reads = int(G * COVERAGE / mean_len)
    lens = _sample_lengths(rng, chrom, n_reads)
    starts = rng.integers(0, G, size=n_reads)
    order = np.argsort(starts, kind="stable")
    starts, ends = starts[order], starts[order] + lens[order]
    # containment removal: drop read j if some earlier-starting read ends at or after end_j
    run_max = np.maximum.accumulate(np.concatenate(([-1], ends[:-1])))
    keep = ends > run_max
    starts, ends = starts[keep], ends[keep]
    R = len(starts)
    read_len_r = (ends - starts).astype(np.int64)
    # + strand overlaps i -> j iff start_i < start_j < end_i - min_overlap
    lo = np.searchsorted(starts, starts, side="right")
    hi = np.searchsorted(starts, ends - min_overlap, side="left")
    cnt = np.maximum(hi - lo, 0)
    ri = np.repeat(np.arange(R), cnt)
    rj = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(lo, cnt)
    ol = (ends[ri] - starts[rj]).astype(np.float64)
    y = np.ones(len(ri), dtype=np.float32)
    # false "repeat" overlaps between random read pairs (label 0)
    n_fp = int(p_fp * len(ri))
    if n_fp:
        fi = rng.integers(0, R, size=n_fp)
        fj = rng.integers(0, R, size=n_fp)
        ok = fi != fj
        fi, fj = fi[ok], fj[ok]
        ri = np.concatenate([ri, fi])
        rj = np.concatenate([rj, fj])
        ol = np.concatenate([ol, rng.uniform(min_overlap, 5000, size=len(fi))])
        y = np.concatenate([y, np.zeros(len(fi), dtype=np.float32)])
    sim = rng.uniform(0.99, 1.0, size=len(ri))
    # Raven read ids are not position sorted: random relabel
    relabel = rng.permutation(R)
    ri, rj = relabel[ri], relabel[rj]
    # both strands: (u+ -> v+) and its mirror (v- -> u-)
    src = np.concatenate([2 * ri, 2 * rj + 1])
    dst = np.concatenate([2 * rj, 2 * ri + 1])
    ol = np.concatenate([ol, ol])
    sim = np.concatenate([sim, sim])
    y = np.concatenate([y, y])
    eorder = np.lexsort((dst, src))           # grouped by src, networkx-style
    src, dst, ol, sim, y = src[eorder], dst[eorder], ol[eorder], sim[eorder], y[eorder]
    N = 2 * R
    ol_z = (ol - ol.mean()) / ol.std(ddof=1)          # torch .std() is unbiased (utils.py:72)
    sim_z = (sim - sim.mean()) / sim.std(ddof=1)
    e = np.stack([ol_z, sim_z], axis=1).astype(np.float32)
    in_deg = np.bincount(dst, minlength=N).astype(np.float32)
    out_deg = np.bincount(src, minlength=N).astype(np.float32)
    pr = pagerank_pe(src, dst, N, pe_dim) if pe_dim > 0 else np.zeros((N, 0), np.float32)   # pe_dim=0: structure only
    pe = np.concatenate([in_deg[:, None], out_deg[:, None], pr], axis=1)
    read_length = np.empty(N, dtype=np.int64)
    read_length[2 * relabel] = read_len_r
    read_length[2 * relabel + 1] = read_len_r
    prefix_length = np.maximum(read_length[src] - ol.astype(np.int64), 1)
    return SynthGraph(src.astype(np.int32), dst.astype(np.int32), N, e, pe.astype(np.float32), y,
                      ol.astype(np.float32), sim.astype(np.float32), read_length, prefix_length)


def make_random_graph(num_nodes, num_edges, seed=0, pe_dim=16, isolated_frac=0.0) -> SynthGraph:
    """Small unstructured multigraph for unit tests (self loops, duplicate edges and isolated /
    zero-in-degree nodes on purpose — the edge cases the aggregation must handle)."""
    rng = np.random.default_rng(seed)
    hi = max(1, int(num_nodes * (1.0 - isolated_frac)))
    src = rng.integers(0, hi, size=num_edges).astype(np.int32)
    dst = rng.integers(0, hi, size=num_edges).astype(np.int32)
    e = rng.standard_normal((num_edges, 2)).astype(np.float32)
    in_deg = np.bincount(dst, minlength=num_nodes).astype(np.float32)
    out_deg = np.bincount(src, minlength=num_nodes).astype(np.float32)
    pr = pagerank_pe(src, dst, num_nodes, pe_dim) if num_edges else np.zeros((num_nodes, pe_dim), np.float32)
    pe = np.concatenate([in_deg[:, None], out_deg[:, None], pr], axis=1).astype(np.float32)
    y = (rng.random(num_edges) < 0.7).astype(np.float32)
    return SynthGraph(src, dst, int(num_nodes), e, pe, y, e[:, 0].copy(), e[:, 1].copy())
