"""Shared test helpers: synthetic data, model construction, model -> oracle parameter mapping."""
import copy
import warnings

import numpy as np
import torch

from oracle import plmc_oracle as O
from projected_lmc_b200 import ProjectedGPModel, gp

KNAMES = {0: "rbf", 1: "matern52", 2: "matern32", 3: "matern12"}

VARIANTS = {
    "PLMC": dict(BDN=False, diagonal_B=False, scalar_B=False),
    "PLMC_fast": dict(BDN=True, diagonal_B=True, scalar_B=True),
    "BDN_full": dict(BDN=True, diagonal_B=False, scalar_B=False),
    "BDN_diag": dict(BDN=True, diagonal_B=True, scalar_B=False),
    "M_diag": dict(BDN=False, diagonal_B=True, scalar_B=False),
    "M_scalar": dict(BDN=False, diagonal_B=True, scalar_B=True),
    "oilmm": dict(BDN=True, diagonal_B=True, scalar_B=True, diagonal_R=True, bulk=False),
    "nonbulk_tri": dict(BDN=False, diagonal_B=False, scalar_B=False, bulk=False),
}


def synth(n, d, p, q, seed=0, ns=0):
    """LMC-style synthetic data (latent smooth functions mixed + noise), standardised per task."""
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n + ns, d, generator=g, dtype=torch.float64) * 2 - 1
    W = torch.randn(d, q, generator=g, dtype=torch.float64) * 2.0
    ph = torch.rand(q, generator=g, dtype=torch.float64) * 6.28
    Fl = torch.sin(X @ W + ph) + 0.3 * torch.cos(2.0 * X @ W)
    Hm = torch.randn(q, p, generator=g, dtype=torch.float64)
    Y = Fl @ Hm + 0.2 * torch.randn(n + ns, p, generator=g, dtype=torch.float64)
    Y = (Y - Y[:n].mean(0)) / Y[:n].std(0)
    return X[:n].contiguous(), Y[:n].contiguous(), X[n:].contiguous(), Y[n:].contiguous()


def make_model(X, Y, q, variant="PLMC", kernel="rbf", outputscales=False, seed=0, perturb=True, **extra):
    torch.manual_seed(seed)
    ktype = gp.kernels.RBFKernel if kernel == "rbf" else gp.kernels.MaternKernel
    kk = {} if kernel in ("rbf", "matern52") else {"nu": {"matern32": 1.5, "matern12": 0.5}[kernel]}
    kw = dict(VARIANTS[variant])
    kw.update(extra)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ProjectedGPModel(X, Y, Y.shape[1], q, mean_type=gp.means.ZeroMean, kernel_type=ktype, ker_kwargs=kk,
                             init_lmc_coeffs=True, outputscales=outputscales, **kw)
    if perturb:  # move every parameter off its symmetric init so all gradient paths are exercised
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for name, prm in m.named_parameters():
                if "Q_plus" in name:
                    continue
                prm.add_(0.1 * torch.randn(prm.shape, generator=g, dtype=prm.dtype))
    return m


def oracle_params(m_cpu) -> O.OracleParams:
    """Effective tensors of a CPU copy of the model, still attached to its raw leaf parameters."""
    base = m_cpu._base_kernel()
    lmc = m_cpu.lmc_coefficients
    comps = None
    cov = m_cpu._covar()
    if hasattr(cov, "kernels") or base.lengthscale_prior is not None or base.active_dims not in (
            None, tuple(range(m_cpu.dim))):
        comps = []
        cm = cov
        for ker in (cm.kernels if hasattr(cm, "kernels") else [cm]):
            b = ker.base_kernel if hasattr(ker, "base_kernel") else ker
            pr = b.lengthscale_prior
            loc = width = None
            if pr is not None:
                loc = pr.loc
                # NormalPrior keeps the standard deviation loc*width, the multivariate one the covariance diag(loc*width)
                width = (pr.scale / pr.loc) if hasattr(pr, "scale") else torch.diagonal(pr.covariance_matrix) / pr.loc
            comps.append(O.OracleComponent(
                dims=list(range(m_cpu.dim)) if b.active_dims is None else list(b.active_dims),
                raw_lengthscale=b.raw_lengthscale,
                raw_outputscale=ker.raw_outputscale if hasattr(ker, "base_kernel") else None,
                prior_loc=loc, prior_width=width))
    p = O.OracleParams(
        raw_lengthscale=base.raw_lengthscale,
        raw_noise=m_cpu.likelihood.noise_covar.raw_noise,
        noise_lower=float(m_cpu.likelihood.noise_covar.raw_noise_constraint.lower_bound),
        kernel=KNAMES[base.kernel_id],
        raw_outputscale=cov.raw_outputscale if (hasattr(cov, "base_kernel") and not hasattr(cov, "kernels")) else None,
        scalar_B=m_cpu.scalar_B, diagonal_B=m_cpu.diagonal_B, eps=m_cpu.eps, components=comps,
        inducing_points=m_cpu._inducing().inducing_points if m_cpu._inducing() is not None else None,
    )
    if lmc.bulk:
        p.H = lmc.H
    else:
        p.Q_plus, p.R = lmc.Q_plus, lmc.R
        p.R_raw_diag = torch.diagonal(lmc.parametrizations.R.original)
    if hasattr(m_cpu, "M"):
        p.M = m_cpu.M
    if hasattr(m_cpu, "log_B_tilde"):
        p.log_B_tilde = m_cpu.log_B_tilde
    if hasattr(m_cpu, "B_tilde_inv_chol"):
        p.B_tilde_inv_chol = m_cpu.B_tilde_inv_chol
    if hasattr(m_cpu, "Y_squared_norm"):
        p.extra["Y_squared_norm"] = m_cpu.Y_squared_norm
    return p


def cpu_copy(m):
    eng = m._engine
    m._engine = None
    cache = m._pred_cache
    m._pred_cache = None
    try:
        c = copy.deepcopy(m).cpu()
    finally:
        m._engine = eng
        m._pred_cache = cache
    return c


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    den = max(b.abs().max().item(), 1e-300)
    return (a - b).abs().max().item() / den
