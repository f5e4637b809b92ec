"""Isolated run of the Gram builder and the backward sweep (for ncu captures / A-B timing).

    python tools/gram_bench.py [n] [d] [q] [kernel_id]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 21
q = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kid = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
X = (torch.rand(n, d, generator=g, dtype=torch.float64) * 2 - 1).to(dev)
ell = torch.full((q, d), 0.7, dtype=torch.float64, device=dev)
noise = torch.full((q,), 0.5, dtype=torch.float64, device=dev)
np_ = ops.npad(n)
Z, zn = ops.scale_inputs(X, ops.col_mean(X), ell, np_)
K = torch.empty((q, np_, np_), dtype=torch.float64, device=dev)
alpha = torch.randn(q, n, dtype=torch.float64, device=dev)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


bytes_ = 8.0 * q * np_ * (np_ + 128) / 2
t = timeit(lambda: ops.gram(Z, zn, kid, None, noise, K, n))
print(f"gram  n={n} d={d} q={q} kid={kid}: {t:.3f} ms  {bytes_ / t / 1e6:.1f} GB/s")
t = timeit(lambda: ops.grad_sweep(K, alpha, Z, zn, ell, kid, None, n))
print(f"sweep n={n} d={d} q={q} kid={kid}: {t:.3f} ms  {bytes_ / t / 1e6:.1f} GB/s")
ops.sweep_debug(True)
t = timeit(lambda: ops.grad_sweep(K, alpha, Z, zn, ell, kid, None, n))
print(f"sweep (direct-difference kernel) n={n} d={d} q={q} kid={kid}: {t:.3f} ms  {bytes_ / t / 1e6:.1f} GB/s")
ops.sweep_debug(False)
