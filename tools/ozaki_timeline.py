"""Per-tile timeline of CTA 0 of the persistent FP64-via-INT8 GEMM (clock64 stamps; plmc_ozaki_debug)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
M, N, K = (int(x) for x in (sys.argv[1:4] or (22272, 640, 640)))
s = int(sys.argv[4]) if len(sys.argv) > 4 else 7
beta = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
dev = torch.device("cuda:0")
A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
C = torch.zeros(M, N, dtype=torch.float64, device=dev)
ws = torch.empty((ops.lib().plmc_ozaki_ws_bytes(M, N, K, s, 0),), dtype=torch.uint8, device=dev)
ops.ozaki_gemm(0, A, B, C, M, N, K, beta=beta, slices=s, ws=ws); torch.cuda.synchronize()
stamps = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
ops.lib().plmc_ozaki_debug(stamps.data_ptr())
ops.ozaki_gemm(0, A, B, C, M, N, K, beta=beta, slices=s, ws=ws); torch.cuda.synchronize()
ops.lib().plmc_ozaki_debug(None)
t = stamps.cpu().view(64, 8)
t0 = int(t[0, 5])
print("tile  copy_first copy_last | mma_start mma_last_issue | acc_full drained c_written   (cycles since first copy)")
for i in range(64):
    if int(t[i, 0]) == 0:
        break
    r = [int(t[i, k]) - t0 for k in (5, 6, 0, 1, 2, 3, 4)]
    print("%4d  %10d %9d | %9d %14d | %8d %7d %9d" % (i, *r))
