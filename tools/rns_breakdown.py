"""One RNS GEMM and one digit-plane GEMM per shape, for an ncu launch list (per-kernel split of the small-K cost):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/rns_breakdown.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
dev = torch.device("cuda:0")
for (M, N, K, layout) in [(22272, 640, 640, 0), (22272, 1408, 1408, 0), (5632, 5632, 5504, 3)]:
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    A = torch.randn((K, M) if a_mc else (M, K), dtype=torch.float64, device=dev)
    B = torch.randn((K, N) if b_nc else (N, K), dtype=torch.float64, device=dev)
    C = torch.zeros(M, N, dtype=torch.float64, device=dev)
    for _ in range(2):
        ops.rns_gemm(layout, A, B, C, M, N, K, beta=1.0, moduli=16)
        ops.rns_gemm(layout, A, B, C, M, N, K, beta=1.0, moduli=16, flags=1)
        ops.ozaki_gemm(layout, A, B, C, M, N, K, beta=1.0, slices=7)
    torch.cuda.synchronize()
