"""Generate golden input/output vectors by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

It puts /root/reference on sys.path, installs oracle/dgl_shim.py as `dgl` (DGL itself is not
installable here, see that file's header), imports the reference's own `models` / `layers`
packages and runs `models.GraphGatedGCNModel.forward` (models/full_graph.py:22-29) — i.e. the real
layers/gated_gcn_full.py:99-157, layers/processor.py:15-20, layers/score_predictor.py:20-25 — plus
the training loss of train.py:211,253-258 and its backward.  Results are stored as small .pt files
next to this script; tests/test_oracle.py pins the oracle to them and tests/test_gpu_parity.py pins
the CUDA path to them.  /root/reference does not exist on the GPU box, hence the committed fixtures.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import dgl_shim  # noqa: E402

dgl = dgl_shim.install()
import models  # noqa: E402  (the reference's package)

from gnnome_assembly_b200.synth import make_assembly_graph, make_random_graph  # noqa: E402


def run_case(name, graph, d, L, batch_norm, seed, state_dict=None, with_grads=True, dtype=torch.float32):
    torch.manual_seed(seed)
    model = models.GraphGatedGCNModel(1, 2, d, 16, L, 64, batch_norm, 16)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    else:
        # default init leaves norm affine at (1, 0) and biases tiny; perturb so every parameter matters
        with torch.no_grad():
            for n_, p in model.named_parameters():
                if "bn_" in n_:
                    p.add_(0.2 * torch.randn_like(p))
    model = model.to(dtype)
    src = torch.from_numpy(graph.src.astype(np.int64))
    dst = torch.from_numpy(graph.dst.astype(np.int64))
    g = dgl.graph((src, dst), num_nodes=graph.num_nodes)
    x = torch.ones(graph.num_nodes, 1, dtype=dtype)
    e = torch.from_numpy(graph.e).to(dtype)
    pe = torch.from_numpy(graph.pe).to(dtype)
    y = torch.from_numpy(graph.y).to(dtype)
    scores = model(g, x, e, pe)                      # the reference forward
    out = {
        "name": name, "d": d, "L": L, "batch_norm": batch_norm,
        "src": graph.src.copy(), "dst": graph.dst.copy(), "num_nodes": graph.num_nodes,
        "e": graph.e.copy(), "pe": graph.pe.copy(), "y": graph.y.copy(),
        "scores": scores.detach().float().clone(),
    }
    if state_dict is None:
        out["state_dict"] = {k: v.detach().float().clone() for k, v in model.state_dict().items()}
    if with_grads:
        pos_weight = torch.tensor([1.0 / 16.5], dtype=dtype)
        loss = torch.nn.BCEWithLogitsLoss(pos_weight=pos_weight)(scores.squeeze(-1), y)
        loss.backward()
        out["pos_weight"] = 1.0 / 16.5
        out["loss"] = float(loss)
        out["grads"] = {k: p.grad.detach().float().clone() for k, p in model.named_parameters()}
    torch.save(out, os.path.join(HERE, f"{name}.pt"))
    print(f"{name}: N={graph.num_nodes} E={graph.num_edges} |scores|max={scores.abs().max():.4f}")


def run_prep_case():
    """The reference's own utils.preprocess_graph / add_positional_encoding / calculate_tfpn (utils.py)."""
    import utils  # noqa: E402  (the reference's module)
    gsyn = make_assembly_graph("chr19", seed=11, genome_len=500_000)
    src = torch.from_numpy(gsyn.src.astype(np.int64))
    dst = torch.from_numpy(gsyn.dst.astype(np.int64))
    g = dgl.graph((src, dst), num_nodes=gsyn.num_nodes)
    g.edata["overlap_length"] = torch.from_numpy(gsyn.overlap_length.astype(np.int64))
    g.edata["overlap_similarity"] = torch.from_numpy(gsyn.overlap_similarity)
    g.edata["y"] = torch.from_numpy(gsyn.y)
    g = utils.preprocess_graph(g, "", 0)
    g = utils.add_positional_encoding(g, 16)
    pe = torch.cat((g.ndata["in_deg"].unsqueeze(1), g.ndata["out_deg"].unsqueeze(1), g.ndata["pe"]), dim=1)  # train.py:249-251
    torch.manual_seed(4)
    scores = torch.randn(gsyn.num_edges) * 2
    scores[:5] = 0.0                                      # sigmoid == 0.5 exactly -> rounds to 0
    tfpn = utils.calculate_tfpn(scores, g.edata["y"])
    pw = 1.0 / 16.5
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([pw]))(scores, g.edata["y"])
    torch.save({"src": gsyn.src.copy(), "dst": gsyn.dst.copy(), "num_nodes": gsyn.num_nodes,
                "overlap_length": gsyn.overlap_length.copy(), "overlap_similarity": gsyn.overlap_similarity.copy(),
                "y": gsyn.y.copy(), "e": g.edata["e"].clone(), "pe": pe.clone(), "scores": scores, "tfpn": tfpn,
                "pos_weight": pw, "loss": float(loss)}, os.path.join(HERE, "ref_prep_small.pt"))
    print(f"ref_prep_small: N={gsyn.num_nodes} E={gsyn.num_edges} tfpn={tfpn} loss={float(loss):.6f}")


if __name__ == "__main__":
    run_prep_case()
    small = make_random_graph(96, 700, seed=3, isolated_frac=0.15)
    run_case("ref_rand_d64_L2_bn", small, 64, 2, True, seed=0)
    run_case("ref_rand_d64_L2_ln", small, 64, 2, False, seed=1)
    asm = make_assembly_graph("chr19", seed=5, genome_len=400_000)
    run_case("ref_asm_d128_L3_bn", asm, 128, 3, True, seed=2)
    # shipped checkpoint (d=256, L=16, BN): outputs only, weights stay in /root/reference
    sd = torch.load("/root/reference/pretrained_models/model_15xchr19.pt", map_location="cpu")
    asm21 = make_assembly_graph("chr21", seed=7, genome_len=300_000)
    run_case("ref_asm_ckpt15xchr19", asm21, 256, 16, True, seed=0, state_dict=sd, with_grads=False)
