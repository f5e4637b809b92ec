"""GPU: the INT8 tensor paths where they can actually fail ("pin the INT8 path").

Fixed-precision operand splitting is accurate NORMWISE (error ~ 2^-bits max|row| max|col| K), not componentwise like
a DGEMM, so rows of L / L^-1 with a wide dynamic range lose bits.  Everything here runs with the DEFAULT thresholds
(min_dim 512) at sizes where the INT8 kernels really execute (npad >= 1024), in all three arithmetic modes:
  * long lengthscales with the noise at its e^-9 floor (cond ~ 1e7..1e9) against the CPU oracle;
  * first-bad-pivot `info` and the jitter retry at n >= 1152, same failing member / jitter / loss in every mode;
  * the reduced-precision inverse (K^-1 feeds only the gradients) against the full-precision one at n = 8192;
  * two engines with different settings on two streams at once (no process-wide state), prediction-cache
    invalidation when the shared workspace is rewritten."""
import math
import warnings

import pytest
import torch

from oracle import plmc_oracle as O
from projected_lmc_b200 import ProjectedLMCmll, gp, ops
from projected_lmc_b200.engine import LatentEngine

from .helpers import cpu_copy, make_model, oracle_params, rel_err, synth

pytestmark = pytest.mark.gpu
MODES = ["fp64", "digits", "rns"]


@pytest.fixture
def gemm_mode(request):
    old = LatentEngine.gemm_mode
    LatentEngine.gemm_mode = request.param
    yield request.param
    LatentEngine.gemm_mode = old


def ill_conditioned_model(n, d, p, q, ell, kernel="rbf", seed=0, noise_thresh=-9.0):
    X, Y, Xs, _ = synth(n, d, p, q, seed=seed, ns=64)
    m = make_model(X, Y, q, variant="PLMC", kernel=kernel, perturb=False, noise_thresh=noise_thresh)
    with torch.no_grad():
        base = m._base_kernel()
        # softplus(raw) = ell * (1 + small per-dimension spread); noise at the floor exp(noise_thresh)
        spread = 1.0 + 0.2 * torch.rand(base.raw_lengthscale.shape, generator=torch.Generator().manual_seed(seed))
        base.raw_lengthscale.copy_(torch.log(torch.expm1(ell * spread)))
        m.likelihood.noise_covar.raw_noise.fill_(-30.0)
    return m, X, Y, Xs


@pytest.mark.parametrize("gemm_mode", MODES, indirect=True)
@pytest.mark.parametrize("n,ell", [(2048, 3.0), (3000, 5.0)])
def test_long_lengthscales_with_noise_at_the_floor_match_the_oracle(gemm_mode, n, ell):
    m, X, Y, Xs = ill_conditioned_model(n, 3, 5, 2, ell, seed=n)
    mc = cpu_copy(m)
    m = m.cuda()
    assert abs(m.projected_noise().min().item() - math.exp(-9.0)) < 1e-12          # really at the floor
    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)                               # no jitter may be needed
        loss = -ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
        loss.backward()
        ref = -O.mll(oracle_params(mc), X, Y)
        ref.backward()
    cond = float(m.kernel_cond().max())
    assert cond > 1e6, cond                                                          # the regime the test is about
    assert abs(loss.item() - ref.item()) <= 1e-8 * abs(ref.item()), (loss.item(), ref.item(), cond)
    refg = dict(mc.named_parameters())
    for name, prm in m.named_parameters():
        if refg[name].grad is not None:
            assert rel_err(prm.grad, refg[name].grad) <= 1e-6, (name, cond)
    m.eval()
    m._engine.predict_inverse = False                  # variance term by triangular solves with L
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pred = m.full_likelihood()(m(Xs.cuda()))
        mean_ref, _, var_ref = O.predict(oracle_params(mc), X, Y, Xs)
        m._engine.predict_inverse = True               # ... and by a triangular multiply with the explicit L^-1
        m._pred_cache = None
        pred_inv = m.full_likelihood()(m(Xs.cuda()))
        assert m._pred_cache["inverted"]
    assert rel_err(pred.mean, mean_ref) <= 1e-6 and rel_err(pred.variance, var_ref) <= 1e-6
    assert rel_err(pred_inv.mean, mean_ref) <= 1e-6 and rel_err(pred_inv.variance, var_ref) <= 1e-6


def _cfg(mode, device, n):
    eng = LatentEngine()
    old = LatentEngine.gemm_mode
    LatentEngine.gemm_mode = mode
    try:
        eng.workspace(torch.device(device), 1, n)
    finally:
        LatentEngine.gemm_mode = old
    return eng


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,bad", [(1152, 700), (2560, 1500), (2560, 2559)])
def test_first_bad_pivot_is_reported_like_cholesky_ex(mode, n, bad):
    """A positive definite matrix whose pivot `bad` is pushed negative: info = bad + 1 (1-based, like LAPACK /
    torch.linalg.cholesky_ex) in every arithmetic mode; a second, healthy batch member reports 0."""
    g = torch.Generator().manual_seed(n + bad)
    A = torch.randn(2, n, n // 4, generator=g, dtype=torch.float64)
    K = (A @ A.transpose(1, 2) / (n // 4) + torch.eye(n, dtype=torch.float64)).cuda()
    Lref = torch.linalg.cholesky(K[0])
    K[0, bad, bad] -= 1.5 * Lref[bad, bad] ** 2            # pivot `bad` becomes -0.5 L_bb^2 < 0, earlier pivots unchanged
    _, info_ref = torch.linalg.cholesky_ex(K)
    assert info_ref.tolist() == [bad + 1, 0]
    eng = _cfg(mode, "cuda", n)
    Kw = K.clone()
    dinv = ops.alloc_dinv(n, 2, K.device)
    info = torch.zeros(2, dtype=torch.int32, device=K.device)
    ops.potrf(Kw, dinv, info, eng.cfg_main)
    assert info.tolist() == [bad + 1, 0]
    assert rel_err(torch.tril(Kw[1]), torch.linalg.cholesky(K[1])) < 1e-12


def test_jitter_retry_agrees_between_arithmetic_modes_at_n_1400():
    """Duplicated inputs and a noise floor of e^-40: latent 0 (long lengthscale) is numerically singular and needs
    jitter, latent 1 (short lengthscale, own noise) does not.  Same failing member, same jitter, same loss."""
    n = 1400
    X, Y, _, _ = synth(n, 2, 4, 2, seed=21)
    X[n // 2:] = X[:n - n // 2]                         # every point twice
    out = {}
    old = LatentEngine.gemm_mode
    try:
        for mode in MODES:
            LatentEngine.gemm_mode = mode
            m = make_model(X, Y, 2, variant="PLMC", kernel="rbf", perturb=False, noise_thresh=-40.0)
            with torch.no_grad():
                m._base_kernel().raw_lengthscale[0].fill_(3.0)
                m._base_kernel().raw_lengthscale[1].fill_(-2.0)
                m.likelihood.noise_covar.raw_noise[0].fill_(-60.0)
                m.likelihood.noise_covar.raw_noise[1].fill_(-1.0)
            m = m.cuda()
            with gp.settings.cholesky_max_tries(8), warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                loss = -ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
                loss.backward()
            jit = m._engine.last_jitter
            assert jit is not None and any("jitter" in str(x.message) for x in w), mode
            out[mode] = (loss.item(), jit.tolist(), {k: p.grad.clone() for k, p in m.named_parameters()})
            del m
    finally:
        LatentEngine.gemm_mode = old
    l0, j0, g0 = out["fp64"]
    assert j0[0] > 0.0 and j0[1] == 0.0                  # only the singular latent received jitter
    for mode in ("digits", "rns"):
        l, j, g = out[mode]
        assert j == j0, (mode, j, j0)                    # same member, same jitter level
        # K + jitter I has cond ~ n / jitter ~ 1e11: the modes agree to what that conditioning allows
        assert abs(l - l0) <= 1e-5 * abs(l0), (mode, l, l0)


def test_reduced_precision_inverse_only_perturbs_gradients_below_1e_minus_9():
    """K^-1 feeds only the gradient sweep: the default 12 moduli (39-42 bits) and 11 against 16 (55 bits) at
    n = 8192; the loss is bit-identical because it comes from L."""
    X, Y, _, _ = synth(8192, 6, 5, 2, seed=31)
    out = {}
    old = LatentEngine.rns_moduli_kinv, LatentEngine.fp64_slices_kinv, LatentEngine.gemm_mode
    try:
        LatentEngine.gemm_mode = "rns"
        assert old[0] == 12                                   # the default under test
        for kinv in (16, 12, 11):
            LatentEngine.rns_moduli_kinv = kinv
            LatentEngine.fp64_slices_kinv = 7 if kinv == 16 else 6
            m = make_model(X, Y, 2, variant="PLMC", kernel="matern52").cuda()
            loss = -ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
            loss.backward()
            out[kinv] = (loss.item(), {k: p.grad.clone() for k, p in m.named_parameters()})
            del m
    finally:
        LatentEngine.rns_moduli_kinv, LatentEngine.fp64_slices_kinv, LatentEngine.gemm_mode = old
    for kinv in (12, 11):
        assert out[16][0] == out[kinv][0]
        for k in out[16][1]:
            assert rel_err(out[kinv][1][k], out[16][1][k]) <= 1e-9, (kinv, k)


def test_two_engines_with_different_settings_on_two_streams():
    """Per-call configuration: a 16-moduli engine and a pure-FP64 engine interleaved on two streams give exactly
    what each gives alone."""
    X, Y, _, _ = synth(2304, 4, 4, 2, seed=41)
    Xg, Yg = X.cuda(), Y.cuda()

    def build(mode):
        old = LatentEngine.gemm_mode
        LatentEngine.gemm_mode = mode
        try:
            m = make_model(X, Y, 2, variant="PLMC", kernel="rbf").cuda()
            m._engine.gemm_mode = mode                    # instance attribute: survives the class default
            return m
        finally:
            LatentEngine.gemm_mode = old

    def run(m):
        for p in m.parameters():
            p.grad = None
        loss = -ProjectedLMCmll(m.likelihood, m)(m(Xg), Yg)
        loss.backward()
        return loss.detach().clone(), [p.grad.clone() for p in m.parameters()]

    ma, mb = build("rns"), build("fp64")
    alone_a, alone_b = run(ma), run(mb)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(sa):
            both_a = run(ma)
        with torch.cuda.stream(sb):
            both_b = run(mb)
    torch.cuda.synchronize()
    assert torch.equal(alone_a[0], both_a[0]) and all(torch.equal(x, y) for x, y in zip(alone_a[1], both_a[1]))
    assert torch.equal(alone_b[0], both_b[0]) and all(torch.equal(x, y) for x, y in zip(alone_b[1], both_b[1]))
    assert ma._engine.emulation_mode() == "rns" and mb._engine.emulation_mode() == "fp64"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_the_second_device():
    X, Y, _, _ = synth(1536, 3, 4, 2, seed=43)
    out = []
    for dev in ("cuda:0", "cuda:1"):
        m = make_model(X, Y, 2, variant="PLMC", kernel="matern52").to(dev)
        loss = -ProjectedLMCmll(m.likelihood, m)(m(X.to(dev)), Y.to(dev))
        loss.backward()
        out.append((loss.item(), [p.grad.cpu() for p in m.parameters()]))
    assert out[0][0] == out[1][0] and all(torch.equal(a, b) for a, b in zip(out[0][1], out[1][1]))


def test_prediction_cache_is_invalidated_when_the_workspace_is_rewritten():
    """compute_loo / kernel_cond / a training-mode step overwrite the engine workspace the cached factor lives in
    (ADVICE r1): the next prediction must re-factorise, not solve against K^-1 or the raw Gram."""
    X, Y, Xs, _ = synth(700, 3, 5, 2, seed=47, ns=33)
    m = make_model(X, Y, 2, variant="PLMC", kernel="matern52").cuda()
    Xg, Yg, Xsg = X.cuda(), Y.cuda(), Xs.cuda()
    m.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p0 = m(Xsg)
        m.compute_loo()
        p1 = m(Xsg)
        m.kernel_cond()
        p2 = m(Xsg)
    m.train()
    loss = -ProjectedLMCmll(m.likelihood, m)(m(Xg), Yg)
    loss.backward()
    m.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p3 = m(Xsg)
    for p in (p1, p2, p3):
        assert torch.equal(p.mean, p0.mean) and torch.equal(p.variance, p0.variance)


def test_prediction_switches_to_the_explicit_inverse_after_half_n_points():
    """engine.predict_inverse = "auto": triangular solves until the points predicted with one factorisation reach
    n/2, then the factor is inverted in place and the variance term becomes a triangular multiply -- same results
    before and after the switch, and a parameter change starts over."""
    X, Y, Xs, _ = synth(900, 3, 5, 2, seed=5, ns=200)
    m = make_model(X, Y, 2, variant="PLMC", kernel="matern52")
    mc = cpu_copy(m)
    m = m.cuda().eval()
    assert m._engine.predict_inverse == "auto"
    Xsg = Xs.cuda()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, var_ref, _ = O.predict(oracle_params(mc), X, Y, Xs)
        v1 = m(Xsg).variance
        assert not m._pred_cache.get("inverted")             # 200 < 450
        v2 = m(Xsg).variance
        assert not m._pred_cache.get("inverted")             # 400 < 450
        v3 = m(Xsg).variance
        assert m._pred_cache["inverted"]                     # 600 >= 450
        v4 = m(Xsg).variance
    assert torch.equal(v1, v2) and torch.equal(v3, v4)
    assert rel_err(v1, var_ref) <= 1e-7 and rel_err(v3, var_ref) <= 1e-7
