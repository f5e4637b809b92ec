"""FP64-via-INT8 GEMM: digit planes (csrc/ozaki.cu) against residue planes (csrc/ozaki2.cu, CTA pairs and single CTA)
at shapes of the C2 recursion, plus the INT8 tensor-pipe peak.  CUDA events, second call timed.
    python tools/rns_bench.py [quick]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import ops

dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
out = {}


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


scratch = torch.zeros(16, dtype=torch.float64, device=dev)
for cg in (1, 2):
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); n_ops = ops.peak_i8(20000, scratch, cg); e1.record(); torch.cuda.synchronize()
        best = max(best, n_ops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    out[f"peak_i8_cta_group_{cg}_tops"] = best
    print(f"peak_i8 cta_group::{cg}: {best:.0f} TOPS", flush=True)

SHAPES = [(8192, 8192, 8192, 0, 0), (22272, 640, 640, 0, 0), (640, 22272, 768, 0, 0), (22272, 1408, 1408, 0, 0),
          (11136, 2816, 2816, 0, 0), (5632, 5632, 5504, 3, 0), (2816, 2816, 2816, 1, 0), (1408, 1408, 1408, 0, 0),
          (11136, 11136, 11136, 0, 1), (22272, 11136, 11136, 3, 0)]
if quick:
    SHAPES = SHAPES[:4]
rows = []
for (M, N, K, layout, lower) in SHAPES:
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    A = torch.randn((K, M) if a_mc else (M, K), dtype=torch.float64, device=dev)
    B = A if lower else torch.randn((K, N) if b_nc else (N, K), dtype=torch.float64, device=dev)
    C = torch.zeros(M, N, dtype=torch.float64, device=dev)
    ws = torch.empty(max(ops.lib().plmc_ozaki_ws_bytes(M, N, K, 7, lower),
                         ops.lib().plmc_rns_ws_bytes(M, N, K, 16, lower, lower)), dtype=torch.uint8, device=dev)
    fl = 2.0 * M * N * K * (0.5 if lower else 1.0)
    row = {"M": M, "N": N, "K": K, "layout": layout, "lower": lower}
    for name, fn in [
        ("digits7", lambda: ops.ozaki_gemm(layout, A, B, C, M, N, K, beta=1.0, lower=bool(lower), slices=7, same_operand=bool(lower), ws=ws)),
        ("digits6", lambda: ops.ozaki_gemm(layout, A, B, C, M, N, K, beta=1.0, lower=bool(lower), slices=6, same_operand=bool(lower), ws=ws)),
        ("rns16", lambda: ops.rns_gemm(layout, A, B, C, M, N, K, beta=1.0, lower=bool(lower), moduli=16, same_operand=bool(lower), ws=ws)),
        ("rns14", lambda: ops.rns_gemm(layout, A, B, C, M, N, K, beta=1.0, lower=bool(lower), moduli=14, same_operand=bool(lower), ws=ws)),
        ("rns16_1cta", lambda: ops.rns_gemm(layout, A, B, C, M, N, K, beta=1.0, lower=bool(lower), moduli=16, same_operand=bool(lower), ws=ws, flags=1)),
    ]:
        ms = timed(fn)
        row[name + "_ms"] = ms
        row[name + "_tflops"] = fl / ms / 1e9
    rows.append(row)
    print(json.dumps(row), flush=True)
    del A, B, C, ws
out["shapes"] = rows
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rns_bench.json", "w"), indent=1)
