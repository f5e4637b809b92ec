"""Known-answer identities for the projected-LMC path.  TEST INFRASTRUCTURE ONLY.

They are independent of gpytorch and of the projection algebra: each one evaluates
the same quantity through the *dense multitask* model that the projected LMC is
mathematically equivalent to (the paper's claim, and the reason the reference's
loss at projected_lmc.py:1178-1241 has its three correction terms):

  KAT-1  n * MLL_projected == log N(vec(Y); 0, sum_l K_l (x) h_l h_l^T + I_n (x) Sigma)
  KAT-2  projected predictive mean / variance == dense multitask GP posterior
  KAT-3  the analytic gradient formulas implemented by the CUDA sweep
         (W = 1/2 (alpha alpha^T - K^-1); tr W; sum W o k; sum A_ij (z_ik - z_jk)^2)
         == autograd through the oracle
"""
from __future__ import annotations

import math

import torch

from . import plmc_oracle as O


def dense_task_covariance(p: O.OracleParams, X: torch.Tensor) -> torch.Tensor:
    """Cov(vec(Y)) of the full LMC, point-major / task-minor ordering, [n p, n p]."""
    n = X.shape[0]
    K = O.gram(p, X, training=False)                       # [q, n, n]
    Ht = O.lmc_coefficients(p)                             # [q, ptasks]
    Sigma = O.task_noise(p)
    C = torch.kron(torch.eye(n, dtype=X.dtype), Sigma)
    for l in range(p.q):
        C = C + torch.kron(K[l], torch.outer(Ht[l], Ht[l]))
    return C


def dense_log_likelihood(p: O.OracleParams, X: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """log N(vec(Y); 0, dense covariance)."""
    C = dense_task_covariance(p, X)
    y = Y.reshape(-1)
    L = torch.linalg.cholesky(C)
    z = torch.linalg.solve_triangular(L, y[:, None], upper=False)[:, 0]
    return -0.5 * (z.pow(2).sum() + 2 * torch.log(torch.diagonal(L)).sum() + y.numel() * math.log(2 * math.pi))


def dense_posterior(p: O.OracleParams, X: torch.Tensor, Y: torch.Tensor, Xs: torch.Tensor):
    """Posterior mean [n*, ptasks] and marginal variance of f (no noise, no eps) of the dense model."""
    n, ns = X.shape[0], Xs.shape[0]
    Ht = O.lmc_coefficients(p)
    ptasks = Ht.shape[1]
    C = dense_task_covariance(p, X)
    Ks = O.gram(p, X, Xs, training=False)                  # [q, n, n*]
    cross = torch.zeros(n * ptasks, ns * ptasks, dtype=X.dtype)
    for l in range(p.q):
        cross = cross + torch.kron(Ks[l], torch.outer(Ht[l], Ht[l]))
    L = torch.linalg.cholesky(C)
    a = torch.cholesky_solve(Y.reshape(-1, 1), L)
    mean = (cross.T @ a).reshape(ns, ptasks)
    V = torch.linalg.solve_triangular(L, cross, upper=False)
    os_ = O.outputscale(p)
    kss = torch.ones(p.q, dtype=X.dtype) if os_ is None else os_
    prior = (kss[:, None] * Ht.pow(2)).sum(0)              # [ptasks]
    var = prior[None, :].expand(ns, -1) - V.pow(2).sum(0).reshape(ns, ptasks)
    return mean, var


def analytic_latent_grads(kind: str, X: torch.Tensor, ell: torch.Tensor, os_, noise: torch.Tensor,
                          TY: torch.Tensor):
    """The closed forms of SURVEY.md 8a row a6 (what csrc/gram.cu grad_sweep_kernel accumulates).

    ell [q, d], os_ [q] | None, noise [q], TY [q, n]  ->  dlp/d(ell, os, noise, TY)."""
    q, d = ell.shape
    n = X.shape[0]
    xm = X - X.mean(0)
    g_ell = torch.zeros(q, d, dtype=X.dtype)
    g_os = torch.zeros(q, dtype=X.dtype)
    g_noise = torch.zeros(q, dtype=X.dtype)
    g_ty = torch.zeros(q, n, dtype=X.dtype)
    for l in range(q):
        z = xm / ell[l]
        diff = z[:, None, :] - z[None, :, :]               # [n, n, d]
        s = diff.pow(2).sum(-1)
        if kind == "rbf":
            k = torch.exp(-0.5 * s)
            dk = -0.5 * k
        else:
            r = s.clamp_min(1e-30).sqrt()
            if kind == "matern52":
                e = torch.exp(-math.sqrt(5) * r)
                k = (1 + math.sqrt(5) * r + 5.0 / 3.0 * r * r) * e
                dk = -(5.0 / 6.0) * (1 + math.sqrt(5) * r) * e
            elif kind == "matern32":
                e = torch.exp(-math.sqrt(3) * r)
                k = (1 + math.sqrt(3) * r) * e
                dk = -1.5 * e
            else:
                k = torch.exp(-r)
                dk = torch.where(s > 1e-30, -0.5 * k / r, torch.zeros_like(k))
        o = 1.0 if os_ is None else os_[l]
        Kf = o * k + noise[l] * torch.eye(n, dtype=X.dtype)
        Kinv = torch.linalg.inv(Kf)
        alpha = Kinv @ TY[l]
        W = 0.5 * (torch.outer(alpha, alpha) - Kinv)
        g_noise[l] = torch.trace(W)
        g_os[l] = (W * k).sum()
        A = W * o * dk
        g_ell[l] = -2.0 / ell[l] * (A[:, :, None] * diff.pow(2)).sum((0, 1))
        g_ty[l] = -alpha
    return g_ell, g_os, g_noise, g_ty
