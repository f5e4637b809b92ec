"""Small end-to-end pass over every kernel in both arithmetic modes (a target for memory checkers where they
are available; on its own it prints the loss and a prediction checksum per mode, which must agree)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ProjectedLMCmll, ops
from projected_lmc_b200.engine import LatentEngine
from tests.helpers import make_model, synth

torch.set_default_dtype(torch.float64)
LatentEngine.fp64_min_dim = 128            # every GEMM of the recursion through the INT8 kernel as well
for layout in range(4):
    M, N, K = 256, 128, 96
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    A = torch.randn((K, M) if a_mc else (M, K), device="cuda")
    B = torch.randn((K, N) if b_nc else (N, K), device="cuda")
    C = torch.randn(M, N, device="cuda")
    ops.ozaki_gemm(layout, A, B, C, M, N, K, alpha=-1.0, beta=1.0)
P = torch.randn(256, 1056, device="cuda")
C = torch.zeros(256, 256, device="cuda")
ops.ozaki_gemm(0, P, P, C, 256, 256, 1056, lower=True, same_operand=True)
X, Y, Xs, _ = synth(700, 3, 5, 2, seed=1, ns=30)
for mode in (7, 0):
    LatentEngine.fp64_slices = mode
    m = make_model(X, Y, 2, variant="PLMC", kernel="matern52").cuda()
    loss = -ProjectedLMCmll(m.likelihood, m)(m(m.train_inputs[0]), m.train_y)
    loss.backward()
    m.eval()
    with torch.no_grad():
        pred = m(Xs.cuda())
    print(mode, loss.item(), pred.mean.sum().item())
torch.cuda.synchronize()
print("done")
