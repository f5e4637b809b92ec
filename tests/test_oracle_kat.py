"""Known-answer tests of the oracle (CPU): dependency-free identities, see oracle/kat.py."""
import math

import numpy as np
import warnings

import pytest
import torch

from oracle import kat
from oracle import plmc_oracle as O

from .helpers import VARIANTS, make_model, oracle_params, rel_err, synth


@pytest.mark.parametrize("variant", ["PLMC", "PLMC_fast", "BDN_diag", "BDN_full", "M_diag", "oilmm", "nonbulk_tri"])
@pytest.mark.parametrize("kernel", ["rbf", "matern52"])
def test_kat1_mll_equals_dense_multitask_loglik(variant, kernel):
    X, Y, _, _ = synth(40, 2, 6, 3, seed=3)
    m = make_model(X, Y, 3, variant=variant, kernel=kernel)
    with torch.no_grad():
        p = oracle_params(m)
        lhs = O.mll(p, X, Y) * X.shape[0]
        rhs = kat.dense_log_likelihood(p, X, Y)
    assert abs(lhs - rhs) <= 1e-10 * abs(rhs), (lhs.item(), rhs.item())


@pytest.mark.parametrize("variant", ["PLMC", "PLMC_fast", "BDN_diag"])
@pytest.mark.parametrize("kernel,os_", [("rbf", False), ("matern52", True)])
def test_kat2_prediction_equals_dense_posterior(variant, kernel, os_):
    X, Y, Xs, _ = synth(36, 2, 5, 2, seed=4, ns=7)
    m = make_model(X, Y, 2, variant=variant, kernel=kernel, outputscales=os_)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p = oracle_params(m)
        mean, var_f, _ = O.predict(p, X, Y, Xs)
        mean_d, var_d = kat.dense_posterior(p, X, Y, Xs)
    assert rel_err(mean, mean_d) < 1e-10
    assert rel_err(var_f - p.eps, var_d) < 1e-9


@pytest.mark.parametrize("kernel", ["rbf", "matern52", "matern32", "matern12"])
@pytest.mark.parametrize("os_", [False, True])
def test_kat3_analytic_gradients_equal_autograd(kernel, os_):
    torch.manual_seed(0)
    n, d, q = 30, 3, 2
    X = torch.rand(n, d) * 2 - 1
    TY = torch.randn(q, n, requires_grad=True)
    ell = (torch.rand(q, d) + 0.5).requires_grad_()
    osv = (torch.rand(q) + 0.5).requires_grad_() if os_ else None
    noise = (torch.rand(q) * 0.2 + 0.05).requires_grad_()
    K = O.base_kernel(kernel, X, X, ell[:, None, :], zero_diag=False)
    if os_:
        K = K * osv[:, None, None]
    K = K + torch.diag_embed(noise[:, None].expand(-1, n))
    lp = O.mvn_log_prob(K, TY).sum()
    lp.backward()
    g_ell, g_os, g_noise, g_ty = kat.analytic_latent_grads(kernel, X, ell.detach(), None if osv is None else osv.detach(),
                                                           noise.detach(), TY.detach())
    # nu = 1/2 is not differentiable at r = 0: autograd through the expansion-based distance picks up
    # O(1e-7) round-off from the diagonal (sqrt of ~1e-15), the closed form treats it as exactly 0
    # (its training-mode Gram diagonal is exp(-sqrt(round-off)) = 1 - O(1e-8) instead of 1).
    tol = 1e-5 if kernel == "matern12" else 1e-10
    assert rel_err(g_ell, ell.grad) < tol
    assert rel_err(g_noise, noise.grad) < tol
    assert rel_err(g_ty, TY.grad) < tol
    if os_:
        assert rel_err(g_os, osv.grad) < tol


def test_psd_safe_cholesky_jitter_only_on_failing_members():
    A = torch.eye(4).repeat(2, 1, 1)
    A[1, 3, 3] = -1e-9                      # second batch member is not PD
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        L = O.psd_safe_cholesky(A, max_tries=3)
    assert len(w) >= 1
    assert torch.allclose(L[0], torch.eye(4))          # untouched member got no jitter
    assert L[1, 3, 3] > 0
    with pytest.raises(RuntimeError):
        bad = torch.eye(3)
        bad[2, 2] = -1.0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            O.psd_safe_cholesky(bad[None], max_tries=3)


# ---------------------------------------------------------------------------------------------------------------
# Independent third-party pins of the gpytorch layer the oracle restates (kernel values and the Gaussian
# log-density): scikit-learn's ARD RBF / Matern(nu) kernels and scipy's multivariate normal.  Neither shares code
# with the oracle or with the shim the reference-run fixtures were generated over.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel,nu", [("rbf", None), ("matern52", 2.5), ("matern32", 1.5), ("matern12", 0.5)])
def test_base_kernel_matches_scikit_learn(kernel, nu):
    from sklearn.gaussian_process.kernels import RBF, Matern

    g = torch.Generator().manual_seed(5)
    X1 = torch.rand(37, 4, generator=g) * 2 - 1
    X2 = torch.rand(23, 4, generator=g) * 2 - 1
    ell = torch.rand(3, 4, generator=g) + 0.3
    K = O.base_kernel(kernel, X1, X2, ell[:, None, :], zero_diag=False)
    Ktrain = O.base_kernel(kernel, X1, X1, ell[:, None, :], zero_diag=True)
    for l in range(3):
        ls = ell[l].numpy()
        sk = RBF(length_scale=ls) if nu is None else Matern(length_scale=ls, nu=nu)
        # the expansion-based squared distance (|a|^2 + |b|^2 - 2ab) carries ~1e-15 absolute round-off; the nu = 1/2
        # kernel exp(-sqrt(s)) turns that into ~3e-8 next to coincident points, everywhere else it stays ~1e-15
        tol = 1e-12
        assert abs(K[l].numpy() - sk(X1.numpy(), X2.numpy())).max() < tol
        ref = sk(X1.numpy())
        tol_train = 1e-7 if kernel == "matern12" else 1e-12
        assert abs(Ktrain[l].numpy() - ref).max() < tol_train


def test_mvn_log_prob_matches_scipy():
    from scipy.stats import multivariate_normal

    g = torch.Generator().manual_seed(6)
    n = 60
    X = torch.rand(n, 3, generator=g) * 2 - 1
    ell = torch.rand(2, 3, generator=g) + 0.5
    noise = torch.tensor([0.05, 0.3])
    K = O.base_kernel("matern52", X, X, ell[:, None, :], zero_diag=True) + torch.diag_embed(noise[:, None].expand(-1, n))
    y = torch.randn(2, n, generator=g)
    lp = O.mvn_log_prob(K, y)
    for l in range(2):
        ref = multivariate_normal(mean=None, cov=K[l].numpy(), allow_singular=False).logpdf(y[l].numpy())
        assert abs(lp[l].item() - ref) <= 1e-10 * abs(ref)


def test_psd_safe_cholesky_jitter_schedule_matches_the_published_one():
    """linear_operator 0.5.0 psd_safe_cholesky: jitter 1e-8 * 10^i (float64) on failing members only, NotPSD after
    max_tries; the factor of a member that did not fail is the plain Cholesky factor."""
    g = torch.Generator().manual_seed(7)
    A = torch.randn(2, 12, 5, generator=g)
    K = A @ A.transpose(1, 2)                       # rank 5 of 12: singular
    K[1] = K[1] + 0.5 * torch.eye(12)               # member 1 is fine
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        L = O.psd_safe_cholesky(K, max_tries=6)
    assert torch.allclose(L[1], torch.linalg.cholesky(K[1]), rtol=0, atol=0)
    assert len(w) >= 1
    rec = L[0] @ L[0].T - K[0]
    off = rec - torch.diag(torch.diagonal(rec))
    assert off.abs().max() < 1e-12
    jit = torch.diagonal(rec)
    assert torch.allclose(jit, jit[0].expand_as(jit), rtol=1e-6, atol=1e-16)
    ratio = math.log10(jit[0].item() / 1e-8)
    assert abs(ratio - round(ratio)) < 1e-3 and 0 <= round(ratio) <= 5


# ---------------------------------------------------------------------------------------------------------------
# handle_covar_ extras (projected_lmc.py:131-181): additive `decomp` kernels and lengthscale priors
# ---------------------------------------------------------------------------------------------------------------
def test_additive_kernel_matches_scikit_learn_sum_kernel():
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern

    g = torch.Generator().manual_seed(11)
    X = torch.rand(31, 4, generator=g) * 2 - 1
    Y = torch.randn(31, 5, generator=g)
    m = make_model(X, Y, 2, variant="PLMC", kernel="matern52", decomp=[[0, 1], [1, 2, 3]])
    p = oracle_params(m)
    assert p.components is not None and len(p.components) == 2
    with torch.no_grad():
        K = O.gram(p, X, training=False)
    for l in range(2):
        ref = 0.0
        for c in p.components:
            ell = O.softplus(c.raw_lengthscale)[l, 0].detach().numpy()
            os_ = float(O.softplus(c.raw_outputscale)[l])
            ref = ref + (ConstantKernel(os_) * Matern(length_scale=ell, nu=2.5))(X[:, c.dims].numpy())
        assert abs(K[l].numpy() - ref).max() < 1e-12


def test_lengthscale_prior_densities_match_scipy_and_count_q_times():
    from scipy.stats import multivariate_normal, norm

    g = torch.Generator().manual_seed(12)
    X = torch.rand(25, 3, generator=g) * 2 - 1
    Y = torch.randn(25, 4, generator=g)
    scales, width = torch.tensor([0.7, 1.3, 0.9]), torch.tensor([0.5, 0.25, 0.4])
    m = make_model(X, Y, 2, variant="PLMC", kernel="rbf", decomp=[[0, 2], [1]], prior_scales=scales, prior_width=width)
    # initialised at the prior mean (:169-176), then perturbed by make_model
    p = oracle_params(m)
    c0, c1 = p.components
    ell0 = O.softplus(c0.raw_lengthscale).detach().numpy()[:, 0]
    ell1 = O.softplus(c1.raw_lengthscale).detach().numpy()[:, 0]
    ref0 = sum(multivariate_normal(mean=scales[[0, 2]].numpy(), cov=np.diag((scales * width)[[0, 2]].numpy())).logpdf(e)
               for e in ell0)
    ref1 = sum(norm(loc=float(scales[1]), scale=float(scales[1] * width[1])).logpdf(e).sum() for e in ell1)
    assert abs(float(O.lengthscale_log_prior(c0)) - ref0) < 1e-12 * max(1.0, abs(ref0))
    assert abs(float(O.lengthscale_log_prior(c1)) - ref1) < 1e-12 * max(1.0, abs(ref1))
    with torch.no_grad():
        with_prior = O.mll(p, X, Y)
        for c in p.components:
            c.prior_loc = None
        without = O.mll(p, X, Y)
    q, n = 2, X.shape[0]
    assert abs(float(with_prior - without) - q * (ref0 + ref1) / n) < 1e-12
    # the product's host-side prior objects give the same densities
    for (_, module, prior, closure), ref in zip(m.named_priors(), (ref0, ref1)):
        assert abs(float(prior.log_prob(closure(module)).sum()) - ref) < 1e-12 * max(1.0, abs(ref))


# ---------------------------------------------------------------------------------------------------------------
# inducing points (ExactGPModel(n_inducing_points=m), projected_lmc.py:302-303 -> gpytorch InducingPointKernel)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel,os_", [("rbf", False), ("matern52", True)])
def test_sgpr_with_inducing_points_on_the_data_is_the_exact_gp(kernel, os_):
    """Q = K_fu K_uu^-1 K_uf equals K when U = X: the SGPR loss, its (then vanishing) trace term and the
    eval-mode prediction must collapse onto the exact model."""
    X, Y, Xs, _ = synth(50, 3, 5, 2, seed=13, ns=8)
    m = make_model(X, Y, 2, variant="PLMC", kernel=kernel, outputscales=os_, n_inducing_points=50)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.covar_module.inducing_points.copy_(X)
        p = oracle_params(m)
        a = O.mll(p, X, Y)
        ma, va, _ = O.predict(p, X, Y, Xs)
        p.inducing_points = None
        b = O.mll(p, X, Y)
        mb, vb, _ = O.predict(p, X, Y, Xs)
    assert abs(a - b) <= 1e-9 * abs(b)
    assert rel_err(ma, mb) < 1e-9 and rel_err(va, vb) < 1e-9


def test_sgpr_objective_is_a_lower_bound_of_the_exact_marginal_likelihood():
    """Titsias (2009): log N(y; 0, Q + s I) - tr(K - Q) / (2 s) <= log N(y; 0, K + s I) for any inducing set."""
    X, Y, _, _ = synth(70, 2, 4, 2, seed=14)
    m = make_model(X, Y, 2, variant="PLMC_fast", kernel="rbf", n_inducing_points=12)
    with torch.no_grad():
        p = oracle_params(m)
        TY = O.project_data(p, Y)
        sparse = O.latent_log_probs(p, X, TY)
        p.inducing_points = None
        exact = O.latent_log_probs(p, X, TY)
    assert (sparse <= exact + 1e-9).all() and (sparse < exact - 1e-3).any()
