"""GPU probe: roofline denominators + correctness/timing of the linear-algebra core.

Run on the GPU box:  python tools/gpu_probe.py [--big]
Writes gpurun_out/probe.json.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


scratch = torch.zeros(16, dtype=torch.float64, device=dev)
# ---- peaks
for threads in (128, 256, 512, 1024):
    for bps in (1, 2):
        blocks = 148 * bps
        fl = [0]
        def run():
            fl[0] = ops.peak_dmma(blocks, threads, 20000, scratch)
        t = timeit(run)
        out[f"dmma_tflops_t{threads}_b{bps}"] = fl[0] / t / 1e12
        def run2():
            fl[0] = ops.peak_dfma(blocks, threads, 20000, scratch)
        t = timeit(run2)
        out[f"dfma_tflops_t{threads}_b{bps}"] = fl[0] / t / 1e12
src = torch.empty(1 << 28, dtype=torch.float64, device=dev).normal_()
dst = torch.empty_like(src)
t = timeit(lambda: ops.peak_copy(src, dst), reps=5)
out["copy_gbs"] = 16 * src.numel() / t / 1e9
t = timeit(lambda: dst.copy_(src), reps=5)
out["torch_copy_gbs"] = 16 * src.numel() / t / 1e9
del src, dst
a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
t = timeit(lambda: torch.matmul(a, b))
out["torch_dgemm_8192_tflops"] = 2 * 8192 ** 3 / t / 1e12
print(json.dumps(out, indent=1), flush=True)

# ---- GEMM correctness (all layouts) and speed
torch.manual_seed(0)
M, N, K = 384, 256, 256
for layout in range(4):
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    A = torch.randn(2, K, M, dtype=torch.float64, device=dev) if a_mc else torch.randn(2, M, K, dtype=torch.float64, device=dev)
    B = torch.randn(2, K, N, dtype=torch.float64, device=dev) if b_nc else torch.randn(2, N, K, dtype=torch.float64, device=dev)
    C = torch.randn(2, M, N, dtype=torch.float64, device=dev)
    opA = A.transpose(1, 2) if a_mc else A
    opB = B if b_nc else B.transpose(1, 2)
    ref = 0.7 * opA @ opB - 1.3 * C
    ops.gemm(layout, A, B, C, M, N, K, alpha=0.7, beta=-1.3)
    err = (C - ref).abs().max().item()
    out[f"gemm_layout{layout}_maxerr"] = err
    print("gemm layout", layout, "err", err, flush=True)

n = 8192
A = torch.randn(1, n, n, dtype=torch.float64, device=dev)
B = torch.randn(1, n, n, dtype=torch.float64, device=dev)
C = torch.zeros(1, n, n, dtype=torch.float64, device=dev)
for layout in range(4):
    t = timeit(lambda: ops.gemm(layout, A, B, C, n, n, n))
    out[f"gemm8192_layout{layout}_tflops"] = 2 * n ** 3 / t / 1e12
t = timeit(lambda: ops.gemm(0, A, A, C, n, n, n, alpha=-1.0, beta=1.0, lower=True))
out["syrk8192_lower_tflops"] = (n * (n + 128)) * n / t / 1e12
print(json.dumps({k: v for k, v in out.items() if "8192" in k}, indent=1), flush=True)
del A, B, C

# ---- potrf / trtri / lauum correctness
def spd(b, n):
    X = torch.randn(b, n, n + 64, dtype=torch.float64, device=dev)
    return X @ X.transpose(1, 2) / n + torch.eye(n, dtype=torch.float64, device=dev)

for n in (128, 384, 1024):
    Kmat = spd(3, n)
    K0 = Kmat.clone()
    dinv = ops.alloc_dinv(n, 3, dev)
    info = torch.zeros(3, dtype=torch.int32, device=dev)
    ops.potrf(Kmat, dinv, info)
    L = torch.tril(Kmat)
    Lref = torch.linalg.cholesky(K0)
    e1 = (L - Lref).abs().max().item()
    y = torch.randn(3, n, dtype=torch.float64, device=dev)
    z, alpha, quad, logdet = ops.solve_logdet(Kmat, dinv, y, n)
    aref = torch.cholesky_solve(y.unsqueeze(-1), Lref).squeeze(-1)
    e2 = (alpha - aref).abs().max().item()
    e3 = (logdet - 2 * torch.log(torch.diagonal(Lref, dim1=1, dim2=2)).sum(-1)).abs().max().item()
    ops.trtri(Kmat, dinv)
    e4 = (torch.tril(Kmat) - torch.linalg.inv(Lref)).abs().max().item()
    ops.lauum(Kmat, dinv)
    e5 = (torch.tril(Kmat) - torch.tril(torch.linalg.inv(K0))).abs().max().item()
    out[f"chol_n{n}"] = dict(potrf=e1, alpha=e2, logdet=e3, trtri=e4, potri=e5, info=info.tolist())
    print(n, out[f"chol_n{n}"], flush=True)

# ---- timing
sizes = [(4096, 4), (8192, 4), (16384, 2)] + ([(32768, 1)] if "--big" in sys.argv else [])
for n, b in sizes:
    Kmat = spd(b, n)
    K0 = Kmat.clone()
    dinv = ops.alloc_dinv(n, b, dev)
    info = torch.zeros(b, dtype=torch.int32, device=dev)
    def f_potrf():
        Kmat.copy_(K0)
        ops.potrf(Kmat, dinv, info)
    tcopy = timeit(lambda: Kmat.copy_(K0))
    t = timeit(f_potrf) - tcopy
    out[f"potrf_n{n}_b{b}_tflops"] = b * n ** 3 / 3 / t / 1e12
    tt = timeit(lambda: torch.linalg.cholesky(K0))
    out[f"torch_potrf_n{n}_b{b}_tflops"] = b * n ** 3 / 3 / tt / 1e12
    def f_potri():
        Kmat.copy_(K0)
        ops.potrf(Kmat, dinv, info)
        ops.potri(Kmat, dinv)
    t2 = timeit(f_potri) - tcopy
    out[f"potrf_potri_n{n}_b{b}_tflops"] = b * n ** 3 / t2 / 1e12
    print(n, b, {k: v for k, v in out.items() if f"_n{n}_b{b}" in k}, flush=True)
    del Kmat, K0, dinv

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe.json", "w") as f:
    json.dump(out, f, indent=1)
print("PROBE DONE")
