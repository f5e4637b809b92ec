"""Bring-up / accuracy / speed probe of the tcgen05 INT8 (Ozaki) FP64 GEMM."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)

def check(layout, M, N, K, s, kind="rand", alpha=1.0, beta=0.0, lower=False, same=False):
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    if kind == "int":      # exactly one 7-bit slice: integers / 128
        gen = lambda *sh: torch.randint(-127, 128, sh, device=dev).double() / 128.0
    else:
        gen = lambda *sh: torch.randn(*sh, dtype=torch.float64, device=dev)
    A = gen(K, M) if a_mc else gen(M, K)
    B = (A if same else (gen(K, N) if b_nc else gen(N, K)))
    C0 = torch.randn(M, N, dtype=torch.float64, device=dev)
    C = C0.clone()
    opA = A.T if a_mc else A
    opB = B if b_nc else B.T
    ref = alpha * opA @ opB + beta * C0
    ops.ozaki_gemm(layout, A, B, C, M, N, K, alpha=alpha, beta=beta, lower=lower, slices=s, same_operand=same)
    torch.cuda.synchronize()
    if lower:
        mask = torch.zeros(M, N, dtype=torch.bool, device=dev)
        for ti in range(M // 128):
            for tj in range(N // 64):
                if not (64 * tj > 128 * ti + 127):
                    mask[128 * ti:128 * ti + 128, 64 * tj:64 * tj + 64] = True
        err = ((C - ref)[mask]).abs().max().item()
        untouched = torch.equal(C[~mask], C0[~mask])
    else:
        err = (C - ref).abs().max().item(); untouched = True
    scale = ref.abs().max().item()
    print(f"layout={layout} M={M} N={N} K={K} s={s} {kind} a={alpha} b={beta} lower={lower} same={same}: "
          f"max|err|={err:.3e} rel={err / scale:.3e} untouched={untouched}", flush=True)
    return err / scale

mode = sys.argv[1] if len(sys.argv) > 1 else "bringup"
if mode == "bringup":
    check(0, 128, 128, 32, 1, "int")
    check(0, 128, 128, 64, 1, "int")
    check(0, 256, 128, 512, 1, "int")
    check(0, 128, 128, 32, 2, "rand")
    check(0, 256, 128, 256, 7, "rand")
    for layout in range(4):
        check(layout, 384, 256, 640, 7, "rand", alpha=-1.0, beta=1.0)
    check(0, 384, 384, 512, 7, "rand", alpha=-1.0, beta=1.0, lower=True, same=True)
    check(3, 384, 384, 512, 7, "rand", alpha=1.0, beta=1.0, lower=True, same=True)
    check(0, 256, 256, 40960, 7, "rand", alpha=1.0, beta=1.0)      # several INT32 K-chunks
    for s in (4, 5, 6, 7):
        check(0, 512, 512, 2048, s, "rand")
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    for s in (7, 6, 5):
        A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
        C = torch.zeros(n, n, dtype=torch.float64, device=dev)
        need = ops.lib().plmc_ozaki_ws_bytes(n, n, n, s, 0)
        ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        ops.ozaki_gemm(0, A, B, C, n, n, n, slices=s, ws=ws); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.ozaki_gemm(0, A, B, C, n, n, n, slices=s, ws=ws); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ref = A @ B.T          # layout 0: B(k, n) stored at B[n, k]
        print(f"n={n} s={s}: {best:.2f} ms  {2 * n ** 3 / best / 1e9:.1f} TFLOP/s (FP64-equivalent)  "
              f"rel err {((C - ref).abs().max() / ref.abs().max()).item():.2e}", flush=True)
