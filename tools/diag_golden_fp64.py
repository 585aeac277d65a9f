"""Diagnostic: for a golden case, the gradient error of the tcgen05 (3xTF32) path, the FFMA path and the committed fp32
golden itself, each against the fp64 oracle on the same weights — tells apart "the kernel is wrong" from "the fp32
reference is as noisy as the kernel" for the heavily cancelled gradients under BatchNorm.
    python tools/diag_golden_fp64.py ref_rand_d64_L2_bn"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gnnome_assembly_b200 as gg
from gnnome_assembly_b200 import _lib
from oracle.gatedgcn_oracle import OracleModel, bce_loss

name = sys.argv[1] if len(sys.argv) > 1 else "ref_rand_d64_L2_bn"
g = torch.load(os.path.join("tests", "golden", f"{name}.pt"), weights_only=False)
dev = torch.device("cuda:0")
src, dst = torch.from_numpy(g["src"].astype(np.int64)), torch.from_numpy(g["dst"].astype(np.int64))


def ours(mode):
    _lib.set_tc_mode(mode)
    m = gg.GraphGatedGCNModel(1, 2, g["d"], 16, g["L"], 64, g["batch_norm"], 16)
    m.load_state_dict(g["state_dict"], strict=True)
    m.to(dev)
    graph = gg.AssemblyGraph(src, dst, g["num_nodes"])
    s = m(graph, None, torch.from_numpy(g["e"]).to(dev), torch.from_numpy(g["pe"]).to(dev))
    bce_loss(s, torch.from_numpy(g["y"]).to(dev), g["pos_weight"]).backward()
    return {k: p.grad.double().cpu() for k, p in m.named_parameters()}


o = OracleModel(1, 2, g["d"], 16, g["L"], 64, g["batch_norm"], 16)
o.load_state_dict(g["state_dict"], strict=True)
o = o.double()
near_zero = {}
for li, conv in enumerate(o.gnn.convs):                       # pre-ReLU values closest to the kink, per layer
    for nm in ("bn_e", "bn_h"):
        def hook(mod, inp, out, key=f"convs.{li}.{nm}"):
            a = out.detach().abs().flatten()
            v, i = torch.sort(a)
            near_zero.setdefault(key, [float(x) for x in v[:3]])
        getattr(conv, nm).register_forward_hook(hook)
r = o(src, dst, g["num_nodes"], torch.from_numpy(g["e"]).double(), torch.from_numpy(g["pe"]).double())
bce_loss(r, torch.from_numpy(g["y"]).double(), g["pos_weight"]).backward()
ref = {k: p.grad for k, p in o.named_parameters()}
tc, ff = ours(1), ours(0)
gold = {k: v.double() for k, v in g["grads"].items()}
print(f"{'tensor':34s} {'max|ref|':>10s} {'tc':>10s} {'ffma':>10s} {'golden32':>10s}")
for k, v in ref.items():
    e = [float((x[k] - v).abs().max()) for x in (tc, ff, gold)]
    print(f"{k:34s} {float(v.abs().max()):10.3e} {e[0]:10.3e} {e[1]:10.3e} {e[2]:10.3e}")
print("smallest |pre-ReLU| per norm (fp64 oracle; an fp32 path flips the mask of anything below ~1e-7):")
for k, v in near_zero.items():
    print(f"  {k:18s} " + " ".join(f"{x:.2e}" for x in v))
