"""ORACLE (test infrastructure only — never imported by the product package).

CPU restatement of the sub-graph step of the reference's mini-batch path: train.py:292-296 / :434-438 draw
clusters with dgl.dataloading.ClusterGCNSampler and train on g.subgraph(node_ids); inference.py:262-275
(get_subgraph) uses dgl.node_subgraph the same way.  DGL is absent here ("parity unpinned" for DGL itself,
DESIGN.md §2); what is restated is DGL's documented contract: sub-graph node j is nodes[j]; the sub-graph's
edges are all parent edges whose two ends are selected, ordered by parent edge id; features are rows of the
parent's; parent ids are kept under dgl.NID / dgl.EID.
"""
import numpy as np


def node_subgraph(src, dst, num_nodes, nodes):
    """-> (sub_src, sub_dst, parent_eid) with DGL's node_subgraph conventions."""
    src, dst, nodes = np.asarray(src), np.asarray(dst), np.asarray(nodes, dtype=np.int64)
    local = np.full(num_nodes, -1, dtype=np.int64)
    local[nodes] = np.arange(nodes.size)
    keep = (local[src] >= 0) & (local[dst] >= 0) if src.size else np.zeros(0, dtype=bool)
    eid = np.nonzero(keep)[0]
    return local[src[eid]], local[dst[eid]], eid


def node_subgraph_loops(src, dst, num_nodes, nodes):
    """the same by plain Python loops (small cases): independent check of the vectorised form"""
    pos = {int(u): j for j, u in enumerate(nodes)}
    s, d, ids = [], [], []
    for i, (a, b) in enumerate(zip(src, dst)):
        if int(a) in pos and int(b) in pos:
            s.append(pos[int(a)]); d.append(pos[int(b)]); ids.append(i)
    return np.array(s, dtype=np.int64), np.array(d, dtype=np.int64), np.array(ids, dtype=np.int64)


def plan_arrays(src, dst, num_nodes, node_perm=None):
    """The engine's plan for (src, dst) in numpy (include/gnnome_b200.h, gg_plan_create): internal node p is
    caller node node_perm[p]; internal edge order = stable sort by internal dst; out-edge CSR over it."""
    src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
    n = int(num_nodes)
    node_perm = np.arange(n) if node_perm is None else np.asarray(node_perm, dtype=np.int64)
    node_inv = np.empty(n, dtype=np.int64)
    node_inv[node_perm] = np.arange(n)
    s, d = node_inv[src], node_inv[dst]
    perm = np.argsort(d, kind="stable")
    inv_perm = np.empty_like(perm)
    inv_perm[perm] = np.arange(perm.size)
    isrc, idst = s[perm], d[perm]
    in_ptr = np.concatenate([[0], np.cumsum(np.bincount(idst, minlength=n))])
    out_ptr = np.concatenate([[0], np.cumsum(np.bincount(isrc, minlength=n))])
    out_eid = np.argsort(isrc, kind="stable")
    return dict(src=isrc, dst=idst, in_ptr=in_ptr, out_ptr=out_ptr, out_eid=out_eid, out_dst=idst[out_eid],
                perm=perm, inv_perm=inv_perm, node_perm=node_perm, node_inv=node_inv)


def subplan_arrays(parent, nodes):
    """Plan of the node-induced sub-graph as gg_subplan_count / gg_subplan_fill define it: the selected nodes keep the parent's
    internal order; everything else follows from plan_arrays on the DGL-convention sub-graph."""
    nodes = np.asarray(nodes, dtype=np.int64)
    n_par = parent["node_perm"].size
    # parent caller edge list
    csrc = parent["node_perm"][parent["src"]][parent["inv_perm"]]
    cdst = parent["node_perm"][parent["dst"]][parent["inv_perm"]]
    s, d, eid = node_subgraph(csrc, cdst, n_par, nodes)
    order = np.argsort(parent["node_inv"][nodes], kind="stable")     # sub-graph node ids by parent-internal position
    out = plan_arrays(s, d, nodes.size, node_perm=order)
    out.update(parent_eid=eid, csrc=s, cdst=d)
    return out
