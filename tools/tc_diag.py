"""GPU diagnostic for the tcgen05 3xTF32 GEMM: each operand-layout variant against fp64, with timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnome_assembly_b200 import _lib
from gnnome_assembly_b200._lib import ptr, check

dev = torch.device("cuda:0")
lib = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (M, N, K) in [(128, 128, 32), (128, 128, 128), (1000, 128, 128), (372650, 128, 128), (45864, 640, 128), (100000, 256, 256)]:
    torch.manual_seed(0)
    x = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    g = torch.randn(M, N, device=dev)
    out = {}
    for mode in (1, 2, 0):
        _lib.set_tc_mode(mode)
        y = torch.empty(M, N, device=dev)
        f1 = lambda: check(lib.gg_linear_fwd(M, N, K, ptr(x), ptr(W), ptr(b), 0, ptr(y), st()), "fwd")
        t1 = timeit(f1)
        gx = torch.empty(M, K, device=dev)
        f2 = lambda: check(lib.gg_linear_bwd_data(M, N, K, ptr(g), ptr(W), None, None, ptr(gx), st()), "bwd_data")
        t2 = timeit(f2)
        dW = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
        f3 = lambda: check(lib.gg_linear_bwd_weight(M, N, K, ptr(g), ptr(x), ptr(dW), ptr(db), st()), "bwd_weight")
        t3 = timeit(f3)
        out[mode] = (y, gx, dW, db, t1, t2, t3)
    y64 = x.double() @ W.double().t() + b.double()
    gx64 = g.double() @ W.double()
    dW64 = g.double().t() @ x.double()
    db64 = g.double().sum(0)
    for mode in (1, 2, 0):
        y, gx, dW, db, t1, t2, t3 = out[mode]
        print(f"M={M} N={N} K={K} mode={['ffma','tc  ','tcraw'][mode]}: NT err {rel(y, y64):.2e} {t1:8.1f}us | "
              f"NN err {rel(gx, gx64):.2e} {t2:8.1f}us | TN err {rel(dW, dW64):.2e} db {rel(db, db64):.2e} {t3:8.1f}us", flush=True)
