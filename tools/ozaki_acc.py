"""Accuracy of the FP64-via-INT8 GEMM per slice count (GPU)."""
import torch
from projected_lmc_b200 import ops

torch.manual_seed(0)
for n, K in ((512, 1024), (1024, 16384), (512, 20480)):
    A = torch.randn(n, K, dtype=torch.float64, device="cuda")
    B = torch.randn(n, K, dtype=torch.float64, device="cuda")
    ref = A @ B.T
    for s in (3, 4, 5, 6, 7):
        C = torch.empty(n, n, dtype=torch.float64, device="cuda")
        ops.ozaki_gemm(0, A, B, C, n, n, K, slices=s)
        print(n, K, s, "max rel-to-max err %.3e" % ((C - ref).abs().max() / ref.abs().max()).item(), flush=True)
# exactness: integers / 64 fit one plane
g = torch.Generator().manual_seed(0)
A = (torch.randint(-63, 64, (256, 512), generator=g).double() / 64).cuda()
B = (torch.randint(-63, 64, (128, 512), generator=g).double() / 64).cuda()
C = torch.empty(256, 128, dtype=torch.float64, device="cuda")
ops.ozaki_gemm(0, A, B, C, 256, 128, 512, slices=1)
print("single-plane exact:", torch.equal(C, A @ B.T))
