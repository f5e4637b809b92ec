"""CPU oracle for the projected-LMC hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference``) may import this module, and only as the checker or the
timed CPU baseline -- never as part of the product path.

What it is: a plain torch (CPU, float64) restatement of the algorithm the
reference executes for ``ProjectedGPModel`` + ``ProjectedLMCmll`` when Cholesky
is forced (``gpytorch.settings.max_cholesky_size`` above n, ``fast_computations``
off).  Each function cites the reference ``file:line`` it follows (paths under the
reference checkout) or, for arithmetic that lives in the un-vendored
dependencies, names the dependency: gpytorch==1.11 / linear_operator==0.5.0
(reference ``requirements.txt:1-2``), whose published algorithm is restated.

PARITY STATUS: **parity unpinned by reference tests** -- the reference ships no
tests, golden vectors or result files for this path, and
gpytorch / linear_operator cannot be installed here.  The oracle is pinned by
(a) golden fixtures generated from the gpytorch-free fragments of the reference
executed in the build container (tests/golden/make_golden.py), (b) a run of the
reference's own ``ProjectedGPModel`` / ``ProjectedLMCmll`` source over a minimal
stand-in for the gpytorch API (oracle/gpytorch_shim, fixtures in tests/golden),
and (c) three dependency-free known-answer identities (oracle/kat.py).
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F

DTYPE = torch.float64


# ---------------------------------------------------------------------------
# parameter container (effective, i.e. already-constrained values are derived
# here from the raw tensors exactly as gpytorch's constraints do)
# ---------------------------------------------------------------------------
@dataclass
class OracleParams:
    """Raw parameters of one ProjectedGPModel (names as in the reference state_dict)."""

    raw_lengthscale: torch.Tensor              # [q, 1, d]   covar_module(.base_kernel).raw_lengthscale
    raw_noise: torch.Tensor                    # [q, 1]      likelihood.noise_covar.raw_noise
    noise_lower: float                         # GreaterThan(exp(noise_thresh)), projected_lmc.py:920-921
    kernel: str = "rbf"                        # rbf | matern52 | matern32 | matern12
    raw_outputscale: Optional[torch.Tensor] = None   # [q] when outputscales=True (ScaleKernel)
    # mixing matrix, bulk mode: H  [p,p] (mode Q_plus) or [p,q] (mode Q)   projected_lmc.py:843-850
    H: Optional[torch.Tensor] = None
    # non-bulk mode: effective Q_plus [p,p]|[p,q] and R [q,q] (after parametrisations), :852-853
    Q_plus: Optional[torch.Tensor] = None
    R: Optional[torch.Tensor] = None
    R_raw_diag: Optional[torch.Tensor] = None  # parametrizations.R.original diagonal (log scale), :1237
    M: Optional[torch.Tensor] = None           # [q, p-q] when BDN=False, :988
    log_B_tilde: Optional[torch.Tensor] = None       # effective [p-q] (diagonal / scalar variants), :975-981
    B_tilde_inv_chol: Optional[torch.Tensor] = None  # effective lower-tri [p-q,p-q] (full variant), :983-984
    scalar_B: bool = False
    diagonal_B: bool = False
    eps: float = 1e-3                          # :900, :993
    extra: dict = field(default_factory=dict)
    # additive `decomp` kernel (handle_covar_, :131-181): one entry per sub-kernel; None = the single ARD kernel above
    components: Optional[list] = None
    # ExactGPModel(n_inducing_points=m) (:302-303): gpytorch InducingPointKernel, inducing points [m, d] shared by
    # the latents; None = exact GP
    inducing_points: Optional[torch.Tensor] = None

    @property
    def q(self) -> int:
        return self.raw_lengthscale.shape[0]


@dataclass
class OracleComponent:
    """One sub-kernel of handle_covar_'s additive kernel (projected_lmc.py:151-167): an ARD kernel on the input
    dimensions ``dims`` wrapped in a ScaleKernel, with an optional lengthscale prior (:135-149)."""

    dims: list
    raw_lengthscale: torch.Tensor              # [q, 1, len(dims)]
    raw_outputscale: Optional[torch.Tensor]    # [q] (always present when there is more than one component)
    prior_loc: Optional[torch.Tensor] = None   # prior mean of the lengthscales (len(dims) values)
    prior_width: Optional[torch.Tensor] = None  # deviation-to-mean ratio


def softplus(x: torch.Tensor) -> torch.Tensor:
    return F.softplus(x)


def lengthscale_log_prior(c: OracleComponent) -> torch.Tensor:
    """log-density of softplus(raw_lengthscale) under the prior handle_covar_ builds (:141-149):
    one dimension -> Normal(loc, loc*width); several -> MultivariateNormal(loc, diag(loc*width)) -- the
    covariance, not the standard deviation, is loc*width there.  Summed over latents."""
    ell = softplus(c.raw_lengthscale)                                   # [q, 1, d_g]
    loc, width = c.prior_loc, c.prior_width
    if len(c.dims) > 1:
        var = loc * width
        lp = -0.5 * ((ell - loc) ** 2 / var).sum(-1) - 0.5 * torch.log(var).sum() - 0.5 * len(c.dims) * math.log(2 * math.pi)
    else:
        sd = loc * width
        lp = -((ell - loc) ** 2) / (2 * sd**2) - torch.log(sd) - 0.5 * math.log(2 * math.pi)
    return lp.sum()


def lengthscale(p: OracleParams) -> torch.Tensor:
    """gpytorch Positive() constraint: softplus(raw).  [q,1,d]"""
    return softplus(p.raw_lengthscale)


def noise(p: OracleParams) -> torch.Tensor:
    """gpytorch GreaterThan(lb): softplus(raw) + lb.  [q]   (projected_noise, projected_lmc.py:996-1000)"""
    return softplus(p.raw_noise).squeeze(-1) + p.noise_lower


def outputscale(p: OracleParams) -> Optional[torch.Tensor]:
    return None if p.raw_outputscale is None else softplus(p.raw_outputscale)


# ---------------------------------------------------------------------------
# gpytorch kernel semantics (gpytorch/kernels/kernel.py: sq_dist / dist;
# rbf_kernel.py; matern_kernel.py) reached from handle_covar_,
# projected_lmc.py:151-167
# ---------------------------------------------------------------------------
def sq_dist(x1: torch.Tensor, x2: torch.Tensor, zero_diag: bool) -> torch.Tensor:
    """Centred quadratic-expansion squared distance, clamp_min 0.

    ``zero_diag`` reproduces gpytorch's ``x1_eq_x2 and not requires_grad`` branch
    (diagonal filled with 0 only when no gradient is needed).
    """
    adj = x1.mean(-2, keepdim=True)
    a = x1 - adj
    b = x2 - adj
    a2 = a.pow(2).sum(-1, keepdim=True)
    b2 = b.pow(2).sum(-1, keepdim=True)
    left = torch.cat([-2.0 * a, a2, torch.ones_like(a2)], dim=-1)
    right = torch.cat([b, torch.ones_like(b2), b2], dim=-1)
    res = left.matmul(right.transpose(-2, -1))
    if zero_diag:
        res = res.clone()
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min(0)


def base_kernel(kind: str, x1: torch.Tensor, x2: torch.Tensor, ell: torch.Tensor, zero_diag: bool) -> torch.Tensor:
    """Batched ARD kernel [q, n1, n2] for inputs [n, d] and lengthscales [q,1,d]."""
    if kind == "rbf":
        z1, z2 = x1.div(ell), x2.div(ell)
        return sq_dist(z1, z2, zero_diag).div(-2).exp()
    nu = {"matern52": 2.5, "matern32": 1.5, "matern12": 0.5}[kind]
    mean = x1.reshape(-1, x1.size(-1)).mean(0)
    z1, z2 = (x1 - mean).div(ell), (x2 - mean).div(ell)
    r = sq_dist(z1, z2, zero_diag).clamp_min(1e-30).sqrt()
    e = torch.exp(-math.sqrt(2 * nu) * r)
    if nu == 0.5:
        c = 1.0
    elif nu == 1.5:
        c = (math.sqrt(3) * r).add(1)
    else:
        c = (math.sqrt(5) * r).add(1).add(5.0 / 3.0 * r**2)
    return c * e


def gram(p: OracleParams, x1: torch.Tensor, x2: Optional[torch.Tensor] = None, training: bool = True) -> torch.Tensor:
    """covar_module(x) of ExactGPModel.forward, projected_lmc.py:1088-1091 (ScaleKernel when outputscales)."""
    same = x2 is None
    x2 = x1 if same else x2
    if p.components is not None:
        # AdditiveKernel of ScaleKernels, each on its own active dimensions (gpytorch slices x[..., active_dims]
        # before the kernel, so the centring of sq_dist is over the selected columns)
        K = 0.0
        for c in p.components:
            Kg = base_kernel(p.kernel, x1[:, c.dims], x2[:, c.dims], softplus(c.raw_lengthscale),
                             zero_diag=(same and not training))
            if c.raw_outputscale is not None:
                Kg = Kg * softplus(c.raw_outputscale)[:, None, None]
            K = K + Kg
        return K
    K = base_kernel(p.kernel, x1, x2, lengthscale(p), zero_diag=(same and not training))
    os_ = outputscale(p)
    if os_ is not None:
        K = K * os_[:, None, None]
    return K


# ---------------------------------------------------------------------------
# linear_operator.utils.cholesky.psd_safe_cholesky
# ---------------------------------------------------------------------------
def psd_safe_cholesky(A: torch.Tensor, max_tries: int = 3, jitter: Optional[float] = None):
    """cholesky_ex; on failure add 1e-8*10^i (f64) / 1e-6*10^i (f32) to the failing batch members only."""
    L, info = torch.linalg.cholesky_ex(A)
    if not torch.any(info):
        return L
    if jitter is None:
        jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
    Aprime = A.clone()
    prev = 0.0
    for i in range(max_tries):
        new = jitter * (10**i)
        diag_add = ((info > 0) * (new - prev)).unsqueeze(-1).expand(*Aprime.shape[:-1])
        Aprime.diagonal(dim1=-1, dim2=-2).add_(diag_add)
        prev = new
        warnings.warn(f"A not p.d., added jitter of {new:.1e} to the diagonal", RuntimeWarning)
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not torch.any(info):
            return L
    raise RuntimeError(f"Matrix not positive definite after repeatedly adding jitter up to {new:.1e}.")


def mvn_log_prob(K: torch.Tensor, y: torch.Tensor, max_tries: int = 3) -> torch.Tensor:
    """gpytorch MultivariateNormal.log_prob with the Cholesky path of inv_quad_logdet (zero mean)."""
    L = psd_safe_cholesky(K, max_tries)
    z = torch.linalg.solve_triangular(L, y.unsqueeze(-1), upper=False).squeeze(-1)
    inv_quad = z.pow(2).sum(-1)
    logdet = 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    n = y.shape[-1]
    return -0.5 * (inv_quad + logdet + n * math.log(2 * math.pi))


# ---------------------------------------------------------------------------
# mixing matrix / projection  (projected_lmc.py:864-884, 1003-1021)
# ---------------------------------------------------------------------------
def qr_factors(p: OracleParams):
    """LMCMixingMatrix.QR, projected_lmc.py:864-875."""
    q = p.q
    if p.H is not None:
        Qp, Rp = torch.linalg.qr(p.H)
        if p.H.shape[0] == p.H.shape[1]:  # mode 'Q_plus'
            return Qp[:, :q], Rp[:q, :q], Qp[:, q:]
        return Qp, Rp, None
    if p.Q_plus.shape[0] == p.Q_plus.shape[1]:
        return p.Q_plus[:, :q], p.R, p.Q_plus[:, q:]
    return p.Q_plus, p.R, None


def lmc_coefficients(p: OracleParams) -> torch.Tensor:
    """LMCMixingMatrix.forward -> H^T [q, p], projected_lmc.py:877-884."""
    if p.H is not None:
        return p.H[:, : p.q].T
    Q, R, _ = qr_factors(p)
    return (Q @ R).T


def projection_matrix(p: OracleParams) -> torch.Tensor:
    """T [p, q], projected_lmc.py:1003-1012."""
    Q, R, Qo = qr_factors(p)
    T = torch.linalg.solve_triangular(R.T, Q, upper=False, left=False)
    if p.M is not None:
        T = T + Qo @ p.M.T * noise(p)[None, :]
    return T


def project_data(p: OracleParams, Y: torch.Tensor) -> torch.Tensor:
    """TY [q, n], projected_lmc.py:1014-1021."""
    Q, R, Qo = qr_factors(p)
    out = torch.linalg.solve_triangular(R, Q.T @ Y.T, upper=True)
    if p.M is not None:
        out = out + noise(p)[:, None] * p.M @ Qo.T @ Y.T
    return out


# ---------------------------------------------------------------------------
# the loss  (ProjectedLMCmll.forward, projected_lmc.py:1178-1241)
# ---------------------------------------------------------------------------
def prior_variance(p: OracleParams, dtype=DTYPE) -> torch.Tensor:
    """k(x, x) of the stationary kernel: the sum of the outputscales (1 without a ScaleKernel).  [q]"""
    if p.components is not None:
        return sum((torch.ones(p.q, dtype=dtype) if c.raw_outputscale is None else softplus(c.raw_outputscale))
                   for c in p.components)
    os_ = outputscale(p)
    return torch.ones(p.q, dtype=dtype) if os_ is None else os_


def sgpr_low_rank(p: OracleParams, x1: torch.Tensor, x2: torch.Tensor, max_tries: int = 3) -> torch.Tensor:
    """gpytorch 1.11 InducingPointKernel._get_covariance without the diagonal correction:
    Q(x1, x2) = K_1u K_uu^-1 K_u2, K_uu factorised by psd_safe_cholesky.  Dense [q, n1, n2]."""
    U = p.inducing_points
    Luu = psd_safe_cholesky(gram(p, U, training=True), max_tries)
    A1 = torch.linalg.solve_triangular(Luu, gram(p, U, x1, training=True), upper=False)
    A2 = A1 if x2 is x1 else torch.linalg.solve_triangular(Luu, gram(p, U, x2, training=True), upper=False)
    return A1.transpose(-1, -2) @ A2


def latent_log_probs(p: OracleParams, X: torch.Tensor, TY: torch.Tensor, max_tries: int = 3) -> torch.Tensor:
    """likelihood(dist).log_prob(TY), projected_lmc.py:1200-1201 -> [q]."""
    n = X.shape[0]
    if p.inducing_points is not None:
        # training-mode InducingPointKernel: covariance Q_ff (no diagonal correction) + noise, and the added loss
        # term -1/2 sum_i (k_ii - q_ii) / noise that ExactMarginalLogLikelihood._add_other_terms adds per latent
        Q = sgpr_low_rank(p, X, X, max_tries)
        s2 = noise(p)
        lp = mvn_log_prob(Q + torch.diag_embed(s2[:, None].expand(-1, n)), TY, max_tries)
        diag = prior_variance(p, X.dtype)[:, None] - torch.diagonal(Q, dim1=-2, dim2=-1)
        return lp - 0.5 * (diag / s2[:, None]).sum(-1)
    K = gram(p, X, training=True)
    n = X.shape[0]
    K = K + torch.diag_embed(noise(p)[:, None].expand(-1, n))
    return mvn_log_prob(K, TY, max_tries)


def projection_terms(p: OracleParams, Y: torch.Tensor):
    """The three correction terms proj_term_list[0..2], projected_lmc.py:1205-1237."""
    n, ptasks = Y.shape
    q = p.q
    Q, R, Qo = qr_factors(p)
    if p.M is None and p.scalar_B:
        if p.log_B_tilde.numel() > 0:
            binv = torch.exp(-p.log_B_tilde[0])
            root_diag = p.log_B_tilde / 2
            ysq = p.extra.get("Y_squared_norm", (Y**2).sum())
            t1 = -0.5 * binv * (ysq - (Y @ Q).pow(2).sum()) / n
        else:
            t1 = torch.zeros((), dtype=Y.dtype)
            root_diag = torch.zeros(1, dtype=Y.dtype)
    else:
        if p.diagonal_B:
            root_diag = p.log_B_tilde / 2
            rot = Y @ Qo
            t1 = -0.5 * torch.trace(rot @ torch.diag_embed(torch.exp(-p.log_B_tilde)) @ rot.T) / n
        else:
            idx = range(ptasks - q)
            root_diag = -torch.log(p.B_tilde_inv_chol[idx, idx])
            root = Y @ Qo @ p.B_tilde_inv_chol
            t1 = -0.5 * torch.trace(root @ root.T) / n
    t0 = -0.5 * 2 * torch.sum(root_diag)
    if p.H is not None:  # bulk
        t2 = -0.5 * torch.log(R[range(q), range(q)] ** 2).sum()
    else:
        t2 = -0.5 * 2 * p.R_raw_diag.sum()
    return t0, t1, t2


def mll(p: OracleParams, X: torch.Tensor, Y: torch.Tensor, max_tries: int = 3) -> torch.Tensor:
    """ProjectedLMCmll.forward, projected_lmc.py:1178-1241 (scalar, per data point)."""
    n, ptasks = Y.shape
    TY = project_data(p, Y)
    lat = latent_log_probs(p, X, TY, max_tries)
    # ExactMarginalLogLikelihood._add_other_terms (:1202): every prior's summed log-density is added to the whole
    # [q] vector (gpytorch 1.11), so it counts q times after the sum below
    for c in (p.components or []):
        if c.prior_loc is not None:
            lat = lat + lengthscale_log_prior(c)
    latent = lat.sum() / n
    t0, t1, t2 = projection_terms(p, Y)
    return latent + (t0 + t1 + t2) - 0.5 * (ptasks - p.q) * math.log(2 * math.pi)


# ---------------------------------------------------------------------------
# prediction  (eval ProjectedGPModel.__call__, projected_lmc.py:1121-1155;
# gpytorch DefaultPredictionStrategy; full_likelihood :1023-1074)
# ---------------------------------------------------------------------------
def task_noise(p: OracleParams) -> torch.Tensor:
    """Sigma [p,p] assembled in full_likelihood, projected_lmc.py:1026-1060."""
    Q, R, Qo = qr_factors(p)
    QR = Q @ R
    sp = noise(p)
    ptasks = Q.shape[0]
    if p.M is not None:
        if p.diagonal_B:
            Broot = torch.diag_embed(torch.exp(p.log_B_tilde / 2))
        else:
            k = p.B_tilde_inv_chol.shape[0]
            Broot = torch.linalg.solve_triangular(p.B_tilde_inv_chol, torch.eye(k, dtype=QR.dtype), upper=False).T
        B = Broot @ Broot.T
        B_term = Qo @ B @ Qo.T
        M_term = -QR @ (sp[:, None] * p.M) @ B @ Qo.T
        D_rot = torch.diag_embed(sp) + sp[:, None] * p.M @ B @ p.M.T * sp[None, :]
        return QR @ D_rot @ QR.T + M_term + M_term.T + B_term
    if p.scalar_B:
        if p.log_B_tilde.numel() > 0:
            B_term = torch.exp(p.log_B_tilde[0]) * (torch.eye(ptasks, dtype=QR.dtype) - Q @ Q.T)
        else:
            B_term = 0.0
    else:
        if p.diagonal_B:
            Broot = torch.diag_embed(torch.exp(p.log_B_tilde / 2))
        else:
            k = p.B_tilde_inv_chol.shape[0]
            Broot = torch.linalg.solve_triangular(p.B_tilde_inv_chol, torch.eye(k, dtype=QR.dtype), upper=False).T
        Br = Qo @ Broot
        B_term = Br @ Br.T
    Dr = QR * torch.sqrt(sp)[None, :]
    return Dr @ Dr.T + B_term


def task_noise_factor(p: OracleParams) -> torch.Tensor:
    """chol(Sigma + eps_c I) with the retry loop of projected_lmc.py:1063-1072 (eps_c from 1e-6, x10 while < eps)."""
    Sigma = task_noise(p).detach()
    ptasks = Sigma.shape[0]
    e = 1e-6
    while e < p.eps:
        try:
            return torch.linalg.cholesky(Sigma + e * torch.eye(ptasks, dtype=Sigma.dtype))
        except Exception:  # noqa: BLE001
            e *= 10
    raise RuntimeError("full noise covariance not PD up to eps (the reference would keep a random factor here)")


def predict(p: OracleParams, X: torch.Tensor, Y: torch.Tensor, Xs: torch.Tensor, max_tries: int = 3):
    """Predictive task means / variances.

    Returns (mean [n*,p], var_f [n*,p], var_y [n*,p]): ``var_f`` is the diagonal of the
    MultitaskMultivariateNormal the model returns (latent posterior mixed through
    H, + eps jitter, projected_lmc.py:1149-1155); ``var_y`` adds diag(F F^T) of the
    full likelihood (:1025, :1068; gpytorch MultitaskGaussianLikelihood, rank=p,
    no global noise).
    """
    n = X.shape[0]
    TY = project_data(p, Y)
    if p.inducing_points is not None:
        # eval-mode InducingPointKernel (sgpr_diagonal_correction on): the joint covariance on [X; X*] is
        # Q + diag(clamp(k_ii - q_ii, 0)); exact-GP prediction under it
        kss_ = prior_variance(p, X.dtype)[:, None]
        Qff = sgpr_low_rank(p, X, X, max_tries)
        corr = (kss_ - torch.diagonal(Qff, dim1=-2, dim2=-1)).clamp_min(0)
        K = Qff + torch.diag_embed(corr + noise(p)[:, None])
        Ks = sgpr_low_rank(p, X, Xs, max_tries)
        qss = torch.diagonal(sgpr_low_rank(p, Xs, Xs, max_tries), dim1=-2, dim2=-1)
        L = psd_safe_cholesky(K, max_tries)
        alpha = torch.cholesky_solve(TY.unsqueeze(-1), L).squeeze(-1)
        lat_mean = (Ks * alpha[:, :, None]).sum(1)
        V = torch.linalg.solve_triangular(L, Ks, upper=False)
        lat_var = qss + (kss_ - qss).clamp_min(0) - V.pow(2).sum(1)
        Ht = lmc_coefficients(p)
        mean = lat_mean.T @ Ht
        var_f = lat_var.T @ Ht.pow(2) + p.eps
        Fch = task_noise_factor(p)
        return mean, var_f, var_f + (Fch @ Fch.T).diagonal()[None, :]
    K = gram(p, X, training=False) + torch.diag_embed(noise(p)[:, None].expand(-1, n))
    L = psd_safe_cholesky(K, max_tries)
    alpha = torch.cholesky_solve(TY.unsqueeze(-1), L).squeeze(-1)          # mean_cache
    Ks = gram(p, X, Xs, training=False)                                     # [q, n, n*]
    lat_mean = (Ks * alpha[:, :, None]).sum(1)                              # [q, n*]
    V = torch.linalg.solve_triangular(L, Ks, upper=False)
    if p.components is not None:
        kss = sum((torch.ones(p.q, dtype=X.dtype) if c.raw_outputscale is None else softplus(c.raw_outputscale))
                  for c in p.components)[:, None]
    else:
        os_ = outputscale(p)
        kss = torch.ones(p.q, 1, dtype=X.dtype) if os_ is None else os_[:, None]
    lat_var = kss - V.pow(2).sum(1)                                         # [q, n*]
    Ht = lmc_coefficients(p)                                                # [q, p]
    mean = lat_mean.T @ Ht
    var_f = lat_var.T @ Ht.pow(2) + p.eps
    Fch = task_noise_factor(p)
    var_y = var_f + (Fch @ Fch.T).diagonal()[None, :]
    return mean, var_f, var_y
