"""torch.autograd glue around the CUDA engine (no arithmetic of its own)."""
from __future__ import annotations

import torch

from . import ops


class ProjectData(torch.autograd.Function):
    """TY = T^T Y^T  (kernel 1).  Gradient flows to T only (Y is data)."""

    @staticmethod
    def forward(ctx, T: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
        T64 = T.detach().to(torch.float64).contiguous()
        ctx.Y = Y
        ctx.in_dtype = T.dtype
        return ops.project_fwd(Y, T64)

    @staticmethod
    def backward(ctx, G: torch.Tensor):
        dT = ops.project_bwd(ctx.Y, G.contiguous())
        return dT.to(ctx.in_dtype), None


class LatentLogProb(torch.autograd.Function):
    """lp_l = log N(TY_l; 0, o_l k_l(X,X) + noise_l I), batched over latents (kernels 2-4).

    The backward quantities are produced eagerly in forward (K -> L -> K^-1 is done in
    place in one cached workspace, so nothing of size n^2 is saved for backward)."""

    @staticmethod
    def forward(ctx, engine, X, kid, TY, ell, os_, noise):
        need = any(t is not None and t.requires_grad for t in (TY, ell, os_, noise))
        lp, grads = engine.log_prob_and_grads(
            X, TY.detach().contiguous(), ell.detach().contiguous(),
            None if os_ is None else os_.detach().contiguous(), noise.detach().contiguous(), kid, need)
        ctx.grads = grads
        ctx.has_os = os_ is not None
        return lp

    @staticmethod
    def backward(ctx, go):
        if ctx.grads is None:
            raise RuntimeError("LatentLogProb.backward called but no input required grad in forward")
        g_ty, g_ell, g_os, g_noise = ctx.grads
        ctx.grads = None
        return (None, None, None, g_ty * go[:, None], g_ell * go[:, None],
                (g_os * go) if ctx.has_os else None, g_noise * go)
