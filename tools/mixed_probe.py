"""Do the FP64 tensor pipe (DMMA) and the FP64 FMA pipe overlap on B200?"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projected_lmc_b200 import _cabi, ops
lib = _cabi.lib()
f = ctypes.CDLL(str(_cabi.lib_path())).plmc_peak_mixed
f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]
scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
def run(blocks, threads, im, iff):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(blocks, threads, im, iff, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
    w = blocks * threads // 64
    mma = w * im * 16 * 512; fma = w * 32 * iff * 16 * 2
    return best, mma, fma
for threads in (256, 512):
    t, m, fl = run(296, threads, 20000, 0);  print(threads, "dmma only  ", round(m / t / 1e12, 2), "TF")
    t, m, fl = run(296, threads, 0, 160000); print(threads, "dfma only  ", round(fl / t / 1e12, 2), "TF")
    for ratio in (4, 8, 16):
        t, m, fl = run(296, threads, 20000, 20000 * ratio)
        print(threads, f"mixed r={ratio}", "dmma", round(m / t / 1e12, 2), "dfma", round(fl / t / 1e12, 2), "sum", round((m + fl) / t / 1e12, 2), "TF")
