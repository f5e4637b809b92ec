"""One triangular product on the residue path for an ncu capture / A-B timing: L^T L (lower) of order n, as one launch
set over the nonzero k-tiles (default) and as the recursion with dense 512-leaves (PLMC_GEMM_FLAG_NO_TRI).
    python tools/tri_one.py [n] [moduli]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mod = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
L0 = torch.tril(torch.randn(1, n, n, dtype=torch.float64, device=dev)) / n ** 0.5
dinv = ops.alloc_dinv(n, 1, dev)
ws = torch.empty(40 << 30, dtype=torch.uint8, device=dev)
for name, flags in (("one launch set", 0), ("recursion", 4)):
    cfg = ops.gemm_cfg(ws, ops.GEMM_INT8_RNS, mod, min_dim=128, flags=flags, alt_precision=6, rns_min_k=1024,
                       rns_min_mnk=int(2e10), min_mnk=512 ** 3)
    best = 1e9
    for _ in range(3):
        L = L0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.lauum(L, dinv, cfg)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"lauum n={n} {mod} moduli, {name}: {best:.2f} ms  {n ** 3 / 3 / best / 1e9:.1f} TFLOP/s-equivalent",
          float(L[0, -1, 0]))
