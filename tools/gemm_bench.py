"""Isolated run of the FP64 DMMA GEMM (for ncu captures and quick A/B timing).

    python tools/gemm_bench.py [n] [layout] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
layout = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
A = torch.randn(1, n, n, dtype=torch.float64, device=dev)
B = torch.randn(1, n, n, dtype=torch.float64, device=dev)
C = torch.zeros(1, n, n, dtype=torch.float64, device=dev)
ops.gemm(layout, A, B, C, n, n, n)
torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(layout, A, B, C, n, n, n)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"gemm n={n} layout={layout}: {best:.3f} ms  {2 * n ** 3 / best / 1e9:.2f} TFLOP/s")
