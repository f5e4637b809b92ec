"""The projected-LMC loss (drop-in for ProjectedLMCmll, projected_lmc.py:1158-1241).

MLL/n = sum_l log N(TY_l | 0, K_l + s_l I)/n  + t0 + t1 + t2 - (p-q)/2 log 2 pi
with the three projection terms kept in ``proj_term_list`` like the reference.  The
latent term runs on the CUDA engine; the O(p^3) projection terms are evaluated from
the p x p second-moment matrix S = Y^T Y (computed once per target with the
projection kernel) instead of the reference's n x n products (:1224, :1229), which
is the same scalar: tr(Q_orth^T S Q_orth B^-1) = tr(Y Q_orth B^-1 Q_orth^T Y^T).
"""
from __future__ import annotations

import numpy as np
import torch

from . import gp, ops


def projection_terms(model, S: torch.Tensor, num_data: int):
    """(t0, t1, t2) of the reference's ``proj_term_list`` given S = Y^T Y  [p, p]."""
    p, q = model.n_tasks, model.n_latents
    Q, R, Q_orth = model.lmc_coefficients.QR()
    if not hasattr(model, 'M') and model.scalar_B:
        if model.log_B_tilde.numel() > 0:
            log_B = model.log_B_tilde
            root_diag = log_B / 2
            y2 = model.Y_squared_norm if hasattr(model, 'Y_squared_norm') else torch.diagonal(S).sum()
            t1 = -0.5 * torch.exp(-log_B[0]) * (y2 - torch.diagonal(Q.T @ S @ Q).sum()) / num_data
        else:
            t1 = 0.
            root_diag = torch.zeros(1, dtype=S.dtype, device=S.device)
    else:
        C = Q_orth.T @ S @ Q_orth                      # (p-q) x (p-q) second moments of the discarded part
        if model.diagonal_B:
            root_diag = model.log_B_tilde / 2
            t1 = -0.5 * (torch.diagonal(C) * torch.exp(-model.log_B_tilde)).sum() / num_data
        else:
            Lb = model.B_tilde_inv_chol
            root_diag = -torch.log(torch.diagonal(Lb))
            t1 = -0.5 * torch.diagonal(Lb.T @ C @ Lb).sum() / num_data
    t0 = -torch.sum(root_diag)
    if model.lmc_coefficients.bulk:
        t2 = -0.5 * torch.log(torch.diagonal(R) ** 2).sum()
    else:
        t2 = -torch.diagonal(model.lmc_coefficients.parametrizations.R.original).sum()
    return t0, t1, t2


class ProjectedLMCmll(gp.mlls.ExactMarginalLogLikelihood):
    """The loss function for the ProjectedGPModel."""

    def __init__(self, latent_likelihood, model):
        super().__init__(latent_likelihood, model)
        self.previous_lat = None
        self._S = None
        self._S_key = None

    def _second_moment(self, target: torch.Tensor) -> torch.Tensor:
        """S = Y^T Y.  Cached only for the model's own ``train_y`` buffer (object identity + version counter);
        any other target tensor is reduced afresh on every call -- a recycled allocation of a different tensor
        can share data_ptr, shape and version with a freed one, so those do not identify it."""
        def second_moment(t):
            Y = t.detach().to(torch.float64).contiguous()
            return ops.project_bwd(Y, Y.T.contiguous())        # S[t, t'] = sum_i Y[i,t] Y[i,t']

        if target is not self.model.train_y:
            return second_moment(target)
        key = (target._version, target.data_ptr(), tuple(target.shape), str(target.device))
        if self._S_key != key:
            self._S = second_moment(target)
            self._S_key = key
        return self._S

    def forward(self, latent_function_dist, target: torch.Tensor, inputs=None, *params):
        if not isinstance(latent_function_dist, gp.distributions.MultivariateNormal):
            raise RuntimeError("ExactMarginalLogLikelihood can only operate on Gaussian random variables")
        model = self.model
        num_data = latent_function_dist.event_shape.numel()
        with model.lmc_coefficients.qr_once():                            # one QR of H for both of its uses
            proj_target = model.project_data(target)                      # n_latents x n_points
            latent_output = self.likelihood(latent_function_dist, *params)
            latent_res = latent_output.log_prob(proj_target)              # latents owned by this process
            latent_res = self._add_other_terms(latent_res, params).sum().div(num_data)

            p, q = model.n_tasks, model.n_latents
            S = self._second_moment(target).to(proj_target.dtype)
            self.proj_term_list = list(projection_terms(model, S, num_data))
        projection_term = sum(self.proj_term_list) - 0.5 * (p - q) * np.log(2 * np.pi)
        world = getattr(model, "_world_size", 1)
        # latent-parallel runs: every rank carries 1/world of the shared terms so that the
        # all-reduced loss and gradients are exactly the single-process ones
        return latent_res + projection_term / world
