"""Golden vectors for the greedy decoder, produced by the reference's OWN code.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_decode.py

Imports /root/reference/inference.py unmodified (oracle/dgl_shim.py stands in for DGL; `graph_dataset` and
`evaluate`, which inference.py imports but get_contigs never calls, are stubbed because they need Biopython)
and runs its get_contigs (inference.py:182-259: get_subgraph, sample_edges, walk_forwards, walk_backwards,
get_contig_length) on a synthetic assembly graph whose successor / predecessor / edge dictionaries come from
the reference's graph_parser functions' logic (edge-id order).  The one random step, sample_edges, runs under
a fixed torch seed and the edges it draws are recorded, so that the oracle and the CUDA decoder can be held
to the same walks bit for bit.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import dgl_shim  # noqa: E402

dgl = dgl_shim.install()
for name in ("graph_dataset", "evaluate"):
    stub = types.ModuleType(name)
    stub.AssemblyGraphDataset = None
    sys.modules[name] = stub
import inference  # noqa: E402  (the reference's module)

from gnnome_assembly_b200.synth import make_assembly_graph  # noqa: E402
from oracle.decode_oracle import adjacency  # noqa: E402


def union(parts):
    """Disjoint union of synthetic graphs (chromosome arms / leftovers); node counts are even, so the
    2k / 2k+1 strand pairing survives the offsets."""
    import dataclasses
    off, fields = 0, {k: [] for k in ("src", "dst", "y", "prefix_length", "overlap_length", "overlap_similarity", "read_length")}
    for p in parts:
        fields["src"].append(p.src.astype(np.int64) + off)
        fields["dst"].append(p.dst.astype(np.int64) + off)
        for k in ("y", "prefix_length", "overlap_length", "overlap_similarity", "read_length"):
            fields[k].append(getattr(p, k))
        off += p.num_nodes
    cat = {k: np.concatenate(v) for k, v in fields.items()}
    return dataclasses.replace(parts[0], num_nodes=off, e=None, pe=None, **cat)


def run_case(name, seed, genome_lens, nb_paths, len_threshold, noise):
    # several disconnected pieces, the last ones shorter than len_threshold reads: the reference's loop only
    # terminates normally (inference.py:243-244) while some edge is left to sample from — on an exhausted graph
    # sample_edges raises (Categorical over zero edges)
    gs = union([make_assembly_graph("chr19", seed=seed + i, genome_len=gl, pe_dim=0) for i, gl in enumerate(genome_lens)])
    src = torch.from_numpy(gs.src.astype(np.int64))
    dst = torch.from_numpy(gs.dst.astype(np.int64))
    g = dgl.graph((src, dst), num_nodes=gs.num_nodes)
    rng = np.random.default_rng(seed)
    # edge logits of a decent but imperfect model: positive for true overlaps, negative for false ones, noisy
    score = ((gs.y * 2 - 1) * 3 + rng.standard_normal(gs.num_edges) * noise).astype(np.float32)
    g.edata["score"] = torch.from_numpy(score)
    g.edata["prefix_length"] = torch.from_numpy(gs.prefix_length)
    g.edata["overlap_length"] = torch.from_numpy(gs.overlap_length)
    g.edata["overlap_similarity"] = torch.from_numpy(gs.overlap_similarity)
    g.ndata["read_length"] = torch.from_numpy(gs.read_length)
    succs, preds, edges = adjacency(gs.src, gs.dst, gs.num_nodes)      # = graph_parser.py:12-73 on g.edges()

    drawn, state = [], {}
    ref_get_subgraph, ref_sample_edges = inference.get_subgraph, inference.sample_edges

    def get_subgraph(g_, visited, device):
        sub_g, m = ref_get_subgraph(g_, visited, device)
        state["sub"], state["map"] = sub_g, m
        return sub_g, m

    def sample_edges(scores_, n):
        idx = ref_sample_edges(scores_, n)
        s, d = state["sub"].edges()
        drawn.append((state["map"][s[idx]].tolist(), state["map"][d[idx]].tolist()))
        return idx

    inference.get_subgraph, inference.sample_edges = get_subgraph, sample_edges
    torch.manual_seed(seed)
    contigs = inference.get_contigs(g, succs, preds, edges, nb_paths=nb_paths, len_threshold=len_threshold, device="cpu")
    # the baselines entry point (inference.py:80-180): model walks + overlap-length / overlap-similarity walks
    drawn_main, drawn_b = list(drawn), []
    drawn.clear()
    base = None
    for attempt in range(20):                             # a draw that exhausts the graph makes the reference raise
        drawn.clear()                                     # (Categorical over zero edges): take the first seed that ends
        torch.manual_seed(seed + 1000 + attempt)          # through the length threshold
        try:
            base = inference.get_contigs_baselines(g, succs, preds, edges, nb_paths=nb_paths, len_threshold=len_threshold, device="cpu")
            break
        except ValueError:
            continue
    assert base is not None
    drawn_b = list(drawn)
    drawn[:] = drawn_main
    inference.get_subgraph, inference.sample_edges = ref_get_subgraph, ref_sample_edges
    torch.save({"baselines": {"start_edges": drawn_b, "contigs": base[0], "contigs_len": base[1], "contigs_sim": base[2]},
                "overlap_length": gs.overlap_length, "overlap_similarity": gs.overlap_similarity,
                "seed": seed, "genome_lens": list(genome_lens), "src": gs.src, "dst": gs.dst,
                "prefix_length": gs.prefix_length, "read_length": gs.read_length, "nb_paths": nb_paths, "len_threshold": len_threshold,
                "score": score, "start_edges": drawn, "contigs": contigs,
                "num_nodes": gs.num_nodes, "num_edges": gs.num_edges}, os.path.join(HERE, name + ".pt"))
    print(f"{name}: N={gs.num_nodes} E={gs.num_edges} iterations={len(drawn)} contigs={[len(c) for c in contigs]}")


if __name__ == "__main__":
    run_case("ref_decode_small", seed=21, genome_lens=(500_000, 250_000, 60_000), nb_paths=8, len_threshold=15, noise=2.5)
    run_case("ref_decode_asm", seed=40, genome_lens=(2_500_000, 1_200_000, 400_000, 80_000), nb_paths=50, len_threshold=20, noise=1.5)
