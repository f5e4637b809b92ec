"""Thin torch-tensor wrappers over the C ABI (one Python function per entry point).

Every wrapper takes CUDA float64 tensors, launches on the current stream and
returns immediately (no synchronisation).  Shapes follow include/plmc_b200.h.
"""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import check, lib, npad, ptr, stream

KERNEL_IDS = {"rbf": 0, "matern52": 1, "matern32": 2, "matern12": 3}

# layout codes of plmc_gemm: bit1 -> A(m,k) m-contiguous, bit0 -> B(k,n) n-contiguous
A_KC_B_KC, A_KC_B_NC, A_MC_B_KC, A_MC_B_NC = 0, 1, 2, 3


def _bstride(t: torch.Tensor) -> int:
    return t.stride(0) if t.dim() == 3 else 0


def gemm(layout, A, B, C, M, N, K, alpha=1.0, beta=0.0, lower=False, triA=False, triB=False):
    """C[b] = alpha * op(A[b]) op(B[b]) + beta * C[b]; A, B, C are [batch, rows, ld] views."""
    batch = C.shape[0]
    check(
        lib().plmc_gemm(
            layout, ptr(A), A.stride(-2), _bstride(A), ptr(B), B.stride(-2), _bstride(B), ptr(C), C.stride(-2),
            _bstride(C), M, N, K, float(alpha), float(beta), int(lower), int(triA), int(triB), batch, stream(),
        ),
        "plmc_gemm",
    )
    return C


def alloc_dinv(np_: int, batch: int, device) -> torch.Tensor:
    return torch.empty((batch, np_ // 128, 128, 128), dtype=torch.float64, device=device)


def potrf(K: torch.Tensor, dinv: torch.Tensor, info: torch.Tensor):
    """In-place lower Cholesky of K [batch, npad, ld]; info int32 [batch]."""
    b, np_, _ = K.shape
    check(lib().plmc_potrf_batched(ptr(K), K.stride(1), K.stride(0), np_, b, ptr(dinv), ptr(info), stream()), "potrf")


def trsm(op: int, L, dinv, B, alpha=1.0):
    """op 0: X L^T = aB | 1: X L = aB (B [batch, m, npad]) | 2: L X = aB | 3: L^T X = aB (B [batch, npad, m])."""
    b, np_, _ = L.shape
    m = B.shape[1] if op in (0, 1) else B.shape[2]
    check(
        lib().plmc_trsm_batched(
            op, ptr(L), L.stride(1), L.stride(0), np_, b, ptr(dinv), ptr(B), B.stride(1), B.stride(0), m,
            float(alpha), stream(),
        ),
        "trsm",
    )


def solve_logdet(L, dinv, y, n, rhs=None):
    """z = L^-1 y, alpha = L^-T z, |z|^2, 2 sum log L_ii for y [batch, >=n]."""
    b, np_, _ = L.shape
    dev = L.device
    if rhs is None:
        rhs = torch.empty((b, np_, 128), dtype=torch.float64, device=dev)
    z = torch.empty((b, n), dtype=torch.float64, device=dev)
    alpha = torch.empty((b, n), dtype=torch.float64, device=dev)
    quad = torch.empty((b,), dtype=torch.float64, device=dev)
    logdet = torch.empty((b,), dtype=torch.float64, device=dev)
    check(
        lib().plmc_solve_logdet(
            ptr(L), L.stride(1), L.stride(0), n, np_, b, ptr(dinv), ptr(y), y.stride(0), ptr(rhs), ptr(z), ptr(alpha),
            n, ptr(quad), ptr(logdet), stream(),
        ),
        "solve_logdet",
    )
    return z, alpha, quad, logdet


def trmv_solve_logdet(Linv, y, n, ws=None):
    """Same outputs as solve_logdet, from the explicit inverse factor Linv [batch, npad, ld]."""
    b, np_, _ = Linv.shape
    dev = Linv.device
    if ws is None:
        ws = torch.empty((b, 128, np_), dtype=torch.float64, device=dev)
    z = torch.empty((b, n), dtype=torch.float64, device=dev)
    alpha = torch.empty((b, n), dtype=torch.float64, device=dev)
    quad = torch.empty((b,), dtype=torch.float64, device=dev)
    logdet = torch.empty((b,), dtype=torch.float64, device=dev)
    check(
        lib().plmc_trmv_solve_logdet(
            ptr(Linv), Linv.stride(1), Linv.stride(0), n, np_, b, ptr(y), y.stride(0), ptr(ws), ptr(z), ptr(alpha), n,
            ptr(quad), ptr(logdet), stream(),
        ),
        "trmv_solve_logdet",
    )
    return z, alpha, quad, logdet


def trtri(L, dinv):
    b, np_, _ = L.shape
    check(lib().plmc_trtri_batched(ptr(L), L.stride(1), L.stride(0), np_, b, ptr(dinv), stream()), "trtri")


def lauum(L):
    b, np_, _ = L.shape
    check(lib().plmc_lauum_batched(ptr(L), L.stride(1), L.stride(0), np_, b, stream()), "lauum")


def potri(L, dinv):
    b, np_, _ = L.shape
    check(lib().plmc_potri_batched(ptr(L), L.stride(1), L.stride(0), np_, b, ptr(dinv), stream()), "potri")


def project_fwd(Y: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    n, p = Y.shape
    q = T.shape[1]
    TY = torch.empty((q, n), dtype=torch.float64, device=Y.device)
    check(lib().plmc_project_fwd(ptr(Y), ptr(T), ptr(TY), n, p, q, n, stream()), "project_fwd")
    return TY


def project_bwd(Y: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    n, p = Y.shape
    q = G.shape[0]
    ws = torch.empty((lib().plmc_project_bwd_ws(n, p, q) // 8,), dtype=torch.float64, device=Y.device)
    dT = torch.empty((p, q), dtype=torch.float64, device=Y.device)
    check(lib().plmc_project_bwd(ptr(Y), ptr(G), G.stride(0), ptr(dT), ptr(ws), n, p, q, stream()), "project_bwd")
    return dT


def dpad_of(d: int) -> int:
    return ((d + 3) // 4) * 4


def col_mean(X: torch.Tensor) -> torch.Tensor:
    n, d = X.shape
    out = torch.empty((d,), dtype=torch.float64, device=X.device)
    check(lib().plmc_col_mean(ptr(X), n, d, ptr(out), stream()), "col_mean")
    return out


def scale_inputs(X, xmean, ell, rows_pad):
    """Z [q, rows_pad, dpad], zn [q, rows_pad]."""
    n, d = X.shape
    q = ell.shape[0]
    dp = dpad_of(d)
    Z = torch.empty((q, rows_pad, dp), dtype=torch.float64, device=X.device)
    zn = torch.empty((q, rows_pad), dtype=torch.float64, device=X.device)
    check(lib().plmc_scale_inputs(ptr(X), ptr(xmean), ptr(ell), ptr(Z), ptr(zn), n, d, dp, rows_pad, q, stream()),
          "scale_inputs")
    return Z, zn


def gram(Z, zn, kernel_id, os_, diag_add, K, n):
    q, np_, dp = Z.shape
    check(
        lib().plmc_gram(ptr(Z), ptr(zn), kernel_id, ptr(os_), ptr(diag_add), ptr(K), K.stride(1), K.stride(0), n, np_,
                        dp, q, stream()),
        "gram",
    )


def cross_gram(Ztr, zntr, Zte, znte, kernel_id, os_, Kx, n, mt):
    q, np_, dp = Ztr.shape
    check(
        lib().plmc_cross_gram(ptr(Ztr), ptr(zntr), ptr(Zte), ptr(znte), kernel_id, ptr(os_), ptr(Kx), Kx.stride(1),
                              Kx.stride(0), n, np_, Zte.shape[1], mt, dp, q, stream()),
        "cross_gram",
    )


def grad_sweep(Kinv, alpha, Z, zn, ell, kernel_id, os_, n):
    q, np_, dp = Z.shape
    d = ell.shape[1]
    dev = Kinv.device
    ws = torch.empty((lib().plmc_grad_ws(np_, d, q) // 8,), dtype=torch.float64, device=dev)
    g_ell = torch.empty((q, d), dtype=torch.float64, device=dev)
    g_os = torch.empty((q,), dtype=torch.float64, device=dev)
    g_noise = torch.empty((q,), dtype=torch.float64, device=dev)
    check(
        lib().plmc_grad_sweep(ptr(Kinv), Kinv.stride(1), Kinv.stride(0), ptr(alpha), alpha.stride(0), ptr(Z), ptr(zn),
                              ptr(ell), kernel_id, ptr(os_), ptr(g_ell), ptr(g_os), ptr(g_noise), ptr(ws), n, np_, d,
                              dp, q, stream()),
        "grad_sweep",
    )
    return g_ell, g_os, g_noise


def latent_mean(Kx, alpha, n, mt):
    q = Kx.shape[0]
    out = torch.empty((q, mt), dtype=torch.float64, device=Kx.device)
    check(lib().plmc_latent_mean(ptr(Kx), Kx.stride(1), Kx.stride(0), ptr(alpha), alpha.stride(0), ptr(out), mt, n, mt,
                                 q, stream()), "latent_mean")
    return out


def latent_var(V, os_, mt):
    q, np_, _ = V.shape
    out = torch.empty((q, mt), dtype=torch.float64, device=V.device)
    check(lib().plmc_latent_var(ptr(V), V.stride(1), V.stride(0), ptr(os_), ptr(out), mt, np_, mt, q, stream()),
          "latent_var")
    return out


def mix_tasks(lat_mean, lat_var, H, var_add, mean, var, mt, accumulate=False):
    q, p = H.shape
    check(lib().plmc_mix_tasks(ptr(lat_mean), ptr(lat_var), lat_mean.stride(0), ptr(H), ptr(var_add), ptr(mean),
                               ptr(var), mt, p, q, int(accumulate), stream()), "mix_tasks")


def peak_dmma(blocks, threads, iters, scratch):
    check(lib().plmc_peak_dmma(blocks, threads, iters, ptr(scratch), stream()), "peak_dmma")
    return blocks * (threads // 32) * iters * 16 * 512


def peak_dfma(blocks, threads, iters, scratch):
    check(lib().plmc_peak_dfma(blocks, threads, iters, ptr(scratch), stream()), "peak_dfma")
    return blocks * threads * iters * 16 * 2


def peak_copy(src, dst):
    n = src.numel()
    check(lib().plmc_peak_copy(ptr(src), ptr(dst), n, stream()), "peak_copy")
    return 16 * n


def stats_reset():
    check(_cabi.load().plmc_stats_reset(), "stats_reset")


def trace_enable(on: bool):
    """Per-shape CUDA-event timing of every GEMM of the factorisation layer (diagnostics)."""
    check(_cabi.load().plmc_trace_enable(int(bool(on))), "trace_enable")


def trace_report():
    check(_cabi.load().plmc_trace_report(), "trace_report")


def stats_get():
    """(kernel launches, GEMM launches, GEMM algorithmic FLOPs) since the last reset."""
    import ctypes

    a, b, c = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_double(0.0)
    check(_cabi.load().plmc_stats_get(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "stats_get")
    return a.value, b.value, c.value


def ozaki_gemm(layout, A, B, C, M, N, K, alpha=1.0, beta=0.0, lower=False, slices=7, same_operand=False, ws=None):
    """C = alpha op(A) op(B) + beta C (2-D views) on the tcgen05 INT8 path (FP64 via 8-bit planes)."""
    need = lib().plmc_ozaki_ws_bytes(M, N, K, slices, int(same_operand))
    if ws is None or ws.numel() < need:
        ws = torch.empty((need,), dtype=torch.uint8, device=C.device)
    check(
        lib().plmc_ozaki_gemm(layout, ptr(A), A.stride(-2), ptr(B), B.stride(-2), ptr(C), C.stride(-2), M, N, K,
                              float(alpha), float(beta), int(lower), slices, int(same_operand), ptr(ws), ws.numel(),
                              stream()),
        "ozaki_gemm",
    )
    return C


_fp64_ws_ref = None   # keeps the scratch the library currently points at alive (the setting is process-wide)


def set_fp64_emulation(ws, slices: int, min_dim: int = 1024):
    """Route the large GEMMs of potrf/trsm/trtri/lauum through the tcgen05 INT8 path (slices=0: off).

    The library stores the raw scratch pointer; this wrapper holds a reference to the tensor for as long as it
    is the configured one, so an engine that goes away cannot leave the library writing into freed memory."""
    global _fp64_ws_ref
    if ws is None or slices == 0:
        check(lib().plmc_set_fp64_emulation(None, 0, 0, max(128, min_dim)), "set_fp64_emulation")
        _fp64_ws_ref = None
    else:
        check(lib().plmc_set_fp64_emulation(ptr(ws), ws.numel(), slices, max(128, min_dim)), "set_fp64_emulation")
        _fp64_ws_ref = ws
