"""FP64-via-INT8 GEMM at the skinny shapes of the recursion (single matrix; events; run under ncu for the
per-kernel split):  python tools/ozaki_shapes.py [slices]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
s = int(sys.argv[1]) if len(sys.argv) > 1 else 7
dev = torch.device("cuda:0")
SHAPES = [(22272, 640, 640, 0), (22272, 640, 640, 3), (640, 22272, 640, 0), (22272, 1408, 1408, 0),
          (11136, 2816, 2816, 0), (5632, 5632, 5504, 0), (2816, 2816, 2816, 0), (1408, 1408, 1408, 0)]
for (M, N, K, layout) in SHAPES:
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    A = torch.randn((K, M) if a_mc else (M, K), dtype=torch.float64, device=dev)
    B = torch.randn((K, N) if b_nc else (N, K), dtype=torch.float64, device=dev)
    C = torch.zeros(M, N, dtype=torch.float64, device=dev)
    ws = torch.empty((ops.lib().plmc_ozaki_ws_bytes(M, N, K, s, 0),), dtype=torch.uint8, device=dev)
    ops.ozaki_gemm(layout, A, B, C, M, N, K, beta=1.0, slices=s, ws=ws); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.ozaki_gemm(layout, A, B, C, M, N, K, beta=1.0, slices=s, ws=ws); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"M={M} N={N} K={K} layout={layout}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
