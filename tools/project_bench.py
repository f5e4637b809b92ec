"""Projection kernels timed alone on a shape that streams from HBM:  python tools/project_bench.py [n] [p] [q]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 32
q = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = "cuda"
Y = torch.randn(n, p, dtype=torch.float64, device=dev)
T = torch.randn(p, q, dtype=torch.float64, device=dev)
G = torch.randn(q, n, dtype=torch.float64, device=dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
b = 8.0 * n * (p + q)
t = timeit(lambda: ops.project_fwd(Y, T)); print(f"fwd n={n} p={p} q={q}: {t:.3f} ms {b / t / 1e6:.0f} GB/s")
t = timeit(lambda: ops.project_bwd(Y, G)); print(f"bwd n={n} p={p} q={q}: {t:.3f} ms {b / t / 1e6:.0f} GB/s")
Yc = torch.empty_like(Y)
t = timeit(lambda: Yc.copy_(Y)); print(f"copy of Y (read + write): {t:.3f} ms {16.0 * n * p / t / 1e6:.0f} GB/s")
