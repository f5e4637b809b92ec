"""One RNS GEMM (default 8192^3, 16 moduli) and one digit-plane GEMM for an ncu --set full capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
A = torch.randn(n, n, dtype=torch.float64, device=dev)
B = torch.randn(n, n, dtype=torch.float64, device=dev)
C = torch.zeros(n, n, dtype=torch.float64, device=dev)
for _ in range(2):
    ops.rns_gemm(0, A, B, C, n, n, n, beta=1.0, moduli=16)
ops.rns_gemm(3, A, B, C, n, n, n, beta=1.0, moduli=16)
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
