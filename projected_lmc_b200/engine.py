"""Device-side orchestration of the latent-GP computations (one instance per model).

This is the Python host side above the C ABI: it owns the (cached) HBM workspaces
and sequences the stream-ordered kernels of csrc/ for
  * training:  scale -> Gram -> potrf (jitter retry) -> solve/logdet -> potri -> fused sweep
  * prediction: factorise once per parameter state, then tile the test points.
It contains no numerical fallback: every array operation on the O(n^2)/O(n^3)
path is a libplmc_b200 kernel.

HBM layout (per model, q_local latents of order n, npad = ceil(n/128)*128):
  K     [q, npad, npad]  f64  Gram -> L -> L^-1 -> K^-1, all in place (lower 128-tiles)
  dinv  [q, npad/128, 128, 128] f64  inverses of the diagonal leaves of L
  rhs   [q, npad, 128]   f64  right-hand-side panel of the two triangular solves
  Z     [q, npad, dpad]  f64  centred inputs divided by the lengthscales; zn [q, npad]
"""
from __future__ import annotations

import math
import warnings

import torch

from . import ops
from ._cabi import npad as _npad
from .gp import settings


class NotPSDError(RuntimeError):
    pass


class NanError(RuntimeError):
    """Raised where linear_operator's psd_safe_cholesky raises its NanError (NaN in the matrix to factor)."""


class LatentEngine:
    def __init__(self):
        self._ws = None
        self._ws_key = None
        self._xmean = None
        self._xmean_key = None
        self.last_jitter = None

    # -- workspaces -----------------------------------------------------------
    def workspace(self, device, q: int, n: int):
        np_ = _npad(n)
        key = (str(device), q, np_)
        if self._ws_key != key:
            self._ws = None  # release before re-allocating
            self._ws = dict(
                K=torch.empty((q, np_, np_), dtype=torch.float64, device=device),
                dinv=ops.alloc_dinv(np_, q, device),
                rhs=torch.empty((q, np_, 128), dtype=torch.float64, device=device),
                info=torch.zeros((q,), dtype=torch.int32, device=device),
            )
            self._ws_key = key
            self._oz = None
        self._configure_fp64(device, np_, q)
        return self._ws

    # -- FP64 through the INT8 tensor path (csrc/ozaki.cu) for the large GEMMs ------------------
    # slices: number of signed 8-bit planes per operand, 8*slices-1 bits (7: DGEMM-grade rounding,
    # 6: 47 bits and ~14 % faster; 0 = pure DMMA arithmetic); min_dim: smallest
    # GEMM dimension routed to the INT8 path.  Defaults come from the environment so that the
    # whole test-suite can be run in either mode.
    fp64_slices = int(__import__("os").environ.get("PLMC_FP64_SLICES", "7"))
    fp64_min_dim = int(__import__("os").environ.get("PLMC_FP64_MIN_DIM", "512"))
    # slices for the explicit inverse of the training iteration (K^-1 = L^-T L^-1: trtri + lauum).  K^-1 feeds ONLY
    # the gradient sweep tr((alpha alpha^T - K^-1) dK): the loss, alpha and the log-determinant come from L, which
    # keeps fp64_slices.  47-bit products there perturb the gradients at the 1e-13 level (tolerance 1e-6) and
    # save a quarter of the time of those two steps; 0 = same as fp64_slices.
    fp64_slices_kinv = int(__import__("os").environ.get("PLMC_FP64_SLICES_KINV", "6"))
    _oz = None

    def _configure_fp64(self, device, np_, q=1):
        s, md = self.fp64_slices, self.fp64_min_dim
        if s <= 0 or np_ < 2 * md:
            ops.set_fp64_emulation(None, 0, md)
            return
        # planes of the largest GEMM of the recursion (s * (M + N) * K bytes, M, K <= npad/2, N <= npad/2
        # or a prediction tile) for every latent, capped: the library works through the batch in passes
        # when the scratch holds fewer members
        per_member = s * (np_ // 2) * (np_ // 2 + 8192) + (1 << 20)
        need = min(q * per_member, max(per_member, 24 << 30))
        if self._oz is None or self._oz.numel() < need or self._oz.device != device:
            self._oz = None
            self._oz = torch.empty((need,), dtype=torch.uint8, device=device)
        ops.set_fp64_emulation(self._oz, s, md)

    def release(self):
        """Drop the HBM workspaces (and un-configure the INT8 path if it points at this engine's scratch)."""
        self._ws, self._ws_key = None, None
        if self._oz is not None:
            if ops._fp64_ws_ref is self._oz:
                ops.set_fp64_emulation(None, 0, self.fp64_min_dim)
            self._oz = None

    def __del__(self):
        try:
            self.release()
        except Exception:   # interpreter shutdown: the library or torch may already be gone
            pass

    def xmean(self, X: torch.Tensor) -> torch.Tensor:
        key = (X.data_ptr(), X._version, tuple(X.shape), str(X.device))
        if self._xmean_key != key:
            self._xmean = ops.col_mean(X)
            self._xmean_key = key
        return self._xmean

    # -- factorisation with gpytorch's psd_safe_cholesky retry semantics ---------
    def _gram_potrf(self, ws, Z, zn, kid, os_, noise, n, max_tries):
        q = Z.shape[0]
        K, dinv, info = ws["K"], ws["dinv"], ws["info"]
        jitter = torch.zeros(q, dtype=torch.float64, device=Z.device)
        # psd_safe_cholesky refuses a matrix with NaN entries before it factors anything.  Every entry of K is a
        # function of zn, Z, the outputscale and the noise, so the O(qn) inputs are checked instead of the n^2
        # matrix (the integer tensor path would turn a NaN into an arbitrary finite number, not propagate it).
        finite = torch.isfinite(zn).all() & torch.isfinite(noise).all()
        if os_ is not None:
            finite = finite & torch.isfinite(os_).all()
        if not bool(finite):
            raise NanError("cholesky: the kernel matrix contains NaN (non-finite inputs, lengthscales, outputscale "
                           "or noise)")
        ops.gram(Z, zn, kid, os_, noise, K, n)
        self._mark("gram")
        ops.potrf(K, dinv, info)
        self._mark("potrf")
        bad = info.cpu()
        if not bool(bad.any()):
            self.last_jitter = None
            return
        base = settings.cholesky_jitter.value()
        prev_bad = bad
        new = 0.0
        for i in range(max_tries):
            new = base * (10**i)
            warnings.warn(f"A not p.d., added jitter of {new:.1e} to the diagonal", RuntimeWarning)
            for l in torch.nonzero(prev_bad).flatten().tolist():
                jitter[l] = new
                sl = slice(l, l + 1)
                da = (noise[sl] + jitter[sl]).contiguous()
                ops.gram(Z[sl], zn[sl], kid, None if os_ is None else os_[sl], da, K[sl], n)
                ops.potrf(K[sl], dinv[sl], info[sl])
            prev_bad = info.cpu()
            if not bool(prev_bad.any()):
                self.last_jitter = jitter
                return
        raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {new:.1e}.")

    # -- training: log-probabilities and all partial gradients -------------------
    def log_prob_and_grads(self, X, TY, ell, os_, noise, kid, need_grad, max_tries=None):
        """lp [q] = log N(TY_l; 0, o_l k_l(X,X) + noise_l I) and, if need_grad,
        (dlp/dTY [q,n], dlp/dell [q,d], dlp/dos [q]|None, dlp/dnoise [q])."""
        if max_tries is None:
            max_tries = settings.cholesky_max_tries.value()
        n, d = X.shape
        q = ell.shape[0]
        ws = self.workspace(X.device, q, n)
        np_ = ws["K"].shape[1]
        mark = self._mark
        mark("start")
        Z, zn = ops.scale_inputs(X, self.xmean(X), ell, np_)
        self._gram_potrf(ws, Z, zn, kid, os_, noise, n, max_tries)
        mark("retry")
        K, dinv = ws["K"], ws["dinv"]
        if not need_grad:
            z, alpha, quad, logdet = ops.solve_logdet(K, dinv, TY, n, ws["rhs"])
            mark("solve_logdet")
            return -0.5 * (quad + logdet + n * math.log(2 * math.pi)), None
        # training step.  z, alpha, the quadratic form and the log-determinant come from L itself (two HBM-bound
        # block substitutions, csrc/trsv.cu) at full FP64-grade accuracy.  The explicit inverse K^-1 = L^-T L^-1
        # (trtri + lauum, 2/3 of the flops of the iteration) then feeds ONLY the gradient sweep
        # tr((alpha alpha^T - K^-1) dK), so both steps may run with fp64_slices_kinv planes.
        z, alpha, quad, logdet = ops.solve_logdet(K, dinv, TY, n, ws["rhs"])
        lp = -0.5 * (quad + logdet + n * math.log(2 * math.pi))
        mark("solve_logdet")
        s_kinv = min(self.fp64_slices_kinv, self.fp64_slices)
        lower = self._oz is not None and 0 < s_kinv < self.fp64_slices and np_ >= 2 * self.fp64_min_dim
        if lower:
            ops.set_fp64_emulation(self._oz, s_kinv, self.fp64_min_dim)
        try:
            ops.trtri(K, dinv)
            ops.lauum(K)
        finally:
            if lower:
                ops.set_fp64_emulation(self._oz, self.fp64_slices, self.fp64_min_dim)
        mark("potri")
        g_ell, g_os, g_noise = ops.grad_sweep(K, alpha, Z, zn, ell, kid, os_, n)
        mark("grad_sweep")
        return lp, (-alpha, g_ell, (g_os if os_ is not None else None), g_noise)

    # -- optional phase timing (CUDA events on the launch stream; used by bench.py) --
    profile = None

    def _mark(self, name):
        if self.profile is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.profile.append((name, ev))

    def phase_ms(self):
        """Sum of the recorded phase durations (ms) by name; call after a synchronize."""
        out = {}
        prev = None
        for name, ev in self.profile or []:
            if name != "start" and prev is not None:
                out[name] = out.get(name, 0.0) + prev.elapsed_time(ev)
            prev = ev
        return out

    # -- dense noisy train covariance (kernel_cond, projected_lmc.py:367-369; small n only) --------
    def dense_gram(self, X, ell, os_, noise, kid):
        n, d = X.shape
        q = ell.shape[0]
        ws = self.workspace(X.device, q, n)
        np_ = ws["K"].shape[1]
        Z, zn = ops.scale_inputs(X, self.xmean(X), ell, np_)
        ops.gram(Z, zn, kid, os_, noise, ws["K"], n)
        K = ws["K"][:, :n, :n]
        return torch.tril(K) + torch.tril(K, -1).transpose(1, 2)

    # -- leave-one-out by-product (projected_lmc.py:1108-1119) ----------------------
    def loo(self, X, TY, ell, os_, noise, kid, max_tries=None):
        if max_tries is None:
            max_tries = settings.cholesky_max_tries.value()
        n, d = X.shape
        q = ell.shape[0]
        ws = self.workspace(X.device, q, n)
        np_ = ws["K"].shape[1]
        Z, zn = ops.scale_inputs(X, self.xmean(X), ell, np_)
        self._gram_potrf(ws, Z, zn, kid, os_, noise, n, max_tries)
        K, dinv = ws["K"], ws["dinv"]
        _, alpha, _, _ = ops.solve_logdet(K, dinv, TY, n, ws["rhs"])
        ops.potri(K, dinv)
        sigma2 = 1.0 / torch.diagonal(K, dim1=1, dim2=2)[:, :n]
        return sigma2, alpha * sigma2

    # -- prediction -----------------------------------------------------------------
    def factorize(self, X, TY, ell, os_, noise, kid, max_tries=None):
        """Prediction cache: L (in ws['K']), dinv, alpha, scaled inputs."""
        if max_tries is None:
            max_tries = settings.cholesky_max_tries.value()
        n, d = X.shape
        q = ell.shape[0]
        ws = self.workspace(X.device, q, n)
        np_ = ws["K"].shape[1]
        xmean = self.xmean(X)
        Z, zn = ops.scale_inputs(X, xmean, ell, np_)
        self._gram_potrf(ws, Z, zn, kid, os_, noise, n, max_tries)
        _, alpha, _, _ = ops.solve_logdet(ws["K"], ws["dinv"], TY, n, ws["rhs"])
        return dict(L=ws["K"], dinv=ws["dinv"], alpha=alpha, Z=Z, zn=zn, xmean=xmean, ell=ell, os=os_, kid=kid, n=n,
                    d=d, q=q)

    @staticmethod
    def tile_points(q: int, np_: int, budget_bytes: int = 6 << 30) -> int:
        mt = (budget_bytes // (q * np_ * 8) // 128) * 128
        return int(max(128, min(8192, mt)))

    def predict_latents(self, st, Xs, need_var=True, tile=None):
        """Latent posterior means / variances at Xs: ([q, n*], [q, n*])."""
        q, n = st["q"], st["n"]
        np_ = st["L"].shape[1]
        ns = Xs.shape[0]
        dev = Xs.device
        self._configure_fp64(dev, np_, q)
        mt_full = tile or self.tile_points(q, np_)
        lat_mean = torch.empty((q, ns), dtype=torch.float64, device=dev)
        lat_var = torch.empty((q, ns), dtype=torch.float64, device=dev) if need_var else None
        Kx = None
        for s0 in range(0, ns, mt_full):
            cnt = min(mt_full, ns - s0)
            mt = _npad(cnt)
            if Kx is None or Kx.shape[2] != mt:
                Kx = None
                Kx = torch.empty((q, np_, mt), dtype=torch.float64, device=dev)
            Zt, znt = ops.scale_inputs(Xs[s0:s0 + cnt].contiguous(), st["xmean"], st["ell"], mt)
            ops.cross_gram(st["Z"], st["zn"], Zt, znt, st["kid"], st["os"], Kx, n, mt)
            lat_mean[:, s0:s0 + cnt] = ops.latent_mean(Kx, st["alpha"], n, mt)[:, :cnt]
            if need_var:
                ops.trsm(2, st["L"], st["dinv"], Kx, 1.0)
                lat_var[:, s0:s0 + cnt] = ops.latent_var(Kx, st["os"], mt)[:, :cnt]
        return lat_mean, lat_var
