"""One large FP64-via-INT8 GEMM (for ncu captures):  python tools/ozaki_one.py [n] [slices]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = torch.device("cuda:0")
A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
C = torch.zeros(n, n, dtype=torch.float64, device=dev)
ws = torch.empty((ops.lib().plmc_ozaki_ws_bytes(n, n, n, s, 0),), dtype=torch.uint8, device=dev)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.ozaki_gemm(0, A, B, C, n, n, n, slices=s, ws=ws); e1.record(); torch.cuda.synchronize()
print(f"n={n} s={s}: {e0.elapsed_time(e1):.2f} ms  {2 * n ** 3 / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s FP64-equivalent")
