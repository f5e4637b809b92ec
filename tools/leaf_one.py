"""One batched 128x128 Cholesky leaf (potrf of order 128, batch 10) for an ncu capture:  python tools/leaf_one.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
b, n = 10, 128
X = torch.randn(b, n, n + 32, dtype=torch.float64, device="cuda")
K0 = X @ X.transpose(1, 2) / n + torch.eye(n, dtype=torch.float64, device="cuda")
dinv = ops.alloc_dinv(n, b, "cuda")
info = torch.zeros(b, dtype=torch.int32, device="cuda")
for _ in range(3):
    K = K0.clone()
    ops.potrf(K, dinv, info)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = K0.clone()
e0.record(); ops.potrf(K, dinv, info); e1.record(); torch.cuda.synchronize()
print("leaf ms", e0.elapsed_time(e1), info.tolist())
