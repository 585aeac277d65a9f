"""ORACLE (test infrastructure only — never imported by the product package).

CPU restatement of the reference's greedy decoder, inference.py:20-77 and :182-259, in plain Python (small
cases).  Pinned by tests/golden/ref_decode_*.pt, which tests/golden/make_golden_decode.py produces by running
the reference's OWN get_contigs / walk_forwards / walk_backwards / get_contig_length (source loaded from
/root/reference/inference.py) on a synthetic assembly graph with a seeded sampler.

The only random step of the reference is sample_edges (inference.py:279-286); it is factored out here as the
`start_edges` argument (per iteration: the (src, dst) pairs the reference drew), so everything below is
deterministic.  arg-max ties: torch.topk(k=1) on the CPU returns the first maximum.
"""
import numpy as np


def greedy_walk(start, scores, adj, edge_id, visited_old, forward=True):
    """inference.py:31-54 (forward=True, adj = successors, edge (current, n)) and :57-77 (forward=False, adj =
    predecessors, edge (n, current), result reversed).  Returns (walk, visited)."""
    walk, seen, cur = [], set(), start
    while True:
        walk.append(cur)
        seen.update((cur, cur ^ 1))                                   # :38-39 / :65-66
        nbrs = adj[cur]
        if len(nbrs) == 0:                                            # :40-41
            break
        if len(nbrs) == 1:                                            # :42-44  (taken without looking at visited)
            cur = nbrs[0]
            continue
        cand = [n for n in nbrs if n not in visited_old and n not in seen]          # :45
        if not cand:                                                  # :47-48
            break
        s = [float(scores[edge_id[(cur, n)] if forward else edge_id[(n, cur)]]) for n in cand]   # :46,49
        cur = cand[int(np.argmax(s))]                                 # :50-51, first maximum
    return (walk if forward else walk[::-1]), seen


def contig_length(walk, prefix_length, read_length, edge_id):
    """inference.py:20-28."""
    total = 0
    for a, b in zip(walk[:-1], walk[1:]):
        total += int(prefix_length[edge_id[(a, b)]])
    return total + int(read_length[walk[-1]])


def walks_for_starts(starts, scores, succs, preds, edge_id, visited):
    """inference.py:118-131: for every sampled edge, forward from its head, backward from its tail."""
    walks, visiteds = [], []
    for s, d in starts:
        wf, vf = greedy_walk(d, scores, succs, edge_id, visited, forward=True)
        wb, vb = greedy_walk(s, scores, preds, edge_id, visited | vf, forward=False)
        walks.append(wb + wf)
        visiteds.append(vf | vb)
    return walks, visiteds


def get_contigs(scores, prefix_length, read_length, succs, preds, edge_id, start_edges, len_threshold=20):
    """inference.py:182-259 with the sampled edges given.  Returns (contigs, visited)."""
    contigs, visited = [], set()
    for starts in start_edges:
        walks, visiteds = walks_for_starts(list(zip(*starts)), scores, succs, preds, edge_id, visited)
        lengths = [contig_length(w, prefix_length, read_length, edge_id) for w in walks]
        best = int(np.argmax(lengths))                                # :221-223 (max -> first maximum)
        walk, seen = walks[best], set(visiteds[best])
        for a, b in zip(walk[:-1], walk[1:]):                         # :227-232 jumped-over nodes and their mates
            t = set(succs[a]) & set(preds[b])
            seen |= t | {x ^ 1 for x in t}
        if len(walk) < len_threshold:                                 # :243-244
            break
        contigs.append(walk)
        visited |= seen                                               # :247
    return contigs, visited


def get_contigs_baselines(score_list, prefix_length, read_length, succs, preds, edge_id, start_edges, len_threshold=20):
    """inference.py:80-180: score_list = [model scores, overlap_length, overlap_similarity]; the model's walks choose
    the contig and the visited set (:146-158,172), the two baselines are walked from the same start edges (:134-141)
    and reported at the chosen index (:175-178).  Returns three lists of walks."""
    out, visited = [[] for _ in score_list], set()
    for starts in start_edges:
        pairs = list(zip(*starts))
        per_score = [walks_for_starts(pairs, sc, succs, preds, edge_id, visited) for sc in score_list]
        walks, visiteds = per_score[0]
        lengths = [contig_length(w, prefix_length, read_length, edge_id) for w in walks]
        best = int(np.argmax(lengths))
        walk, seen = walks[best], set(visiteds[best])
        for a, b in zip(walk[:-1], walk[1:]):
            t = set(succs[a]) & set(preds[b])
            seen |= t | {x ^ 1 for x in t}
        if len(walk) < len_threshold:
            break
        for k in range(len(score_list)):
            out[k].append(per_score[k][0][best])
        visited |= seen
    return out


def adjacency(src, dst, num_nodes):
    """graph_parser.py:12-73: successor / predecessor lists and the edge dictionary, filled in edge-id order."""
    succs = {i: [] for i in range(num_nodes)}
    preds = {i: [] for i in range(num_nodes)}
    edge_id = {}
    for i, (a, b) in enumerate(zip(np.asarray(src).tolist(), np.asarray(dst).tolist())):
        succs[a].append(b)
        preds[b].append(a)
        edge_id[(a, b)] = i
    return succs, preds, edge_id


def edge_weights(src, dst, scores, visited):
    """inference.py:262-286 on the remaining graph: unnormalised sampling weights (fp32 like the reference)."""
    s = np.asarray(scores, dtype=np.float32)
    p = (1.0 / (1.0 + np.exp(-s.astype(np.float64)))).astype(np.float32)
    p = np.maximum(p, np.float32(1e-9))
    vis = np.zeros(int(max(np.max(src, initial=0), np.max(dst, initial=0))) + 2, dtype=bool)
    vis[list(visited)] = True
    p[vis[np.asarray(src)] | vis[np.asarray(dst)] | (np.asarray(src) == np.asarray(dst))] = 0
    return p
