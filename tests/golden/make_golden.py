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


if __name__ == "__main__":
    small = make_random_graph(96, 700, seed=3, isolated_frac=0.15)
    run_case("ref_rand_d64_L2_bn", small, 64, 2, True, seed=0)
    run_case("ref_rand_d64_L2_ln", small, 64, 2, False, seed=1)
    asm = make_assembly_graph("chr19", seed=5, genome_len=400_000)
    run_case("ref_asm_d128_L3_bn", asm, 128, 3, True, seed=2)
    # shipped checkpoint (d=256, L=16, BN): outputs only, weights stay in /root/reference
    sd = torch.load("/root/reference/pretrained_models/model_15xchr19.pt", map_location="cpu")
    asm21 = make_assembly_graph("chr21", seed=7, genome_len=300_000)
    run_case("ref_asm_ckpt15xchr19", asm21, 256, 16, True, seed=0, state_dict=sd, with_grads=False)
