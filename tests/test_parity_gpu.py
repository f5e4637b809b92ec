"""GPU parity: the CUDA path (through the model API and the C ABI) against the CPU oracle.

Tolerances (north_star, fp64): MLL 1e-8 relative, gradients / predictive means / variances
1e-6 relative.  The tests assert a tighter 1e-9 / 1e-7 so regressions show early."""
import warnings

import pytest
import torch

from oracle import plmc_oracle as O
from projected_lmc_b200 import ProjectedLMCmll, gp

from .helpers import VARIANTS, cpu_copy, make_model, oracle_params, rel_err, synth

pytestmark = pytest.mark.gpu

MLL_TOL = 1e-9
GRAD_TOL = 1e-7
PRED_TOL = 1e-7


def run_case(n, d, p, q, variant, kernel, outputscales=False, ns=37, seed=0, grad_tol=GRAD_TOL, mll_tol=MLL_TOL,
             **model_kwargs):
    X, Y, Xs, _ = synth(n, d, p, q, seed=seed, ns=ns)
    m = make_model(X, Y, q, variant=variant, kernel=kernel, outputscales=outputscales, seed=seed, **model_kwargs)
    mc = cpu_copy(m)
    m = m.cuda()
    Xg, Yg = X.cuda(), Y.cuda()
    m.train()
    mll = ProjectedLMCmll(m.likelihood, m)
    loss = -mll(m(Xg), Yg)
    loss.backward()

    mc.train()
    op = oracle_params(mc)
    ref = -O.mll(op, X, Y)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= mll_tol * abs(ref.item()), (loss.item(), ref.item())
    ref_grads = dict(mc.named_parameters())
    for name, prm in m.named_parameters():
        g_ref = ref_grads[name].grad
        if g_ref is None:
            assert prm.grad is None or prm.grad.abs().max().item() == 0.0, name
            continue
        assert prm.grad is not None, name
        assert rel_err(prm.grad, g_ref) <= grad_tol, (name, prm.grad.cpu(), g_ref)

    # prediction
    m.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fl = m.full_likelihood()
        lat = m(Xs.cuda())
        obs = fl(lat)
        mean_ref, varf_ref, vary_ref = O.predict(oracle_params(mc), X, Y, Xs)
    assert rel_err(lat.mean, mean_ref) <= PRED_TOL
    assert rel_err(lat.variance, varf_ref) <= PRED_TOL
    assert rel_err(obs.variance, vary_ref) <= PRED_TOL
    lo, hi = obs.confidence_region()
    assert torch.allclose(hi - lo, 4 * obs.stddev)
    return m


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("kernel", ["rbf", "matern52"])
def test_variants_small(variant, kernel):
    run_case(n=150, d=3, p=7, q=3, variant=variant, kernel=kernel)


@pytest.mark.parametrize("kernel", ["matern32", "matern12"])
def test_other_matern(kernel):
    # nu = 1/2 is not differentiable at r = 0: the reference's training-mode Gram diagonal is
    # exp(-sqrt(round-off)) = 1 - O(1e-8) and its autograd carries the same noise; the CUDA path uses exactly 1
    loose = kernel == "matern12"
    run_case(n=90, d=2, p=5, q=2, variant="PLMC", kernel=kernel, grad_tol=(1e-5 if loose else GRAD_TOL),
             mll_tol=(1e-6 if loose else MLL_TOL))


def test_outputscales():
    m = run_case(n=140, d=4, p=6, q=2, variant="PLMC", kernel="matern52", outputscales=True)
    assert m.outputscale().shape == (2, 1)


@pytest.mark.parametrize("kernel", ["rbf", "matern52"])
def test_additive_decomp_kernels(kernel):
    """handle_covar_ `decomp` (projected_lmc.py:151-167): k = sum_g o_g k_g(x[dims_g]), overlapping groups, every
    sub-kernel with its own ARD lengthscales and outputscale per latent."""
    m = run_case(300, 5, 6, 3, "PLMC", kernel, decomp=[[0, 1], [1, 2, 3], [4]], seed=5)
    ls, os_ = m.lscales(), m.outputscale()
    assert isinstance(ls, list) and [tuple(t.shape) for t in ls] == [(3, 2), (3, 3), (3,)]
    assert tuple(os_.shape) == (3, 3)
    assert any(name.startswith("covar_module.kernels.2.base_kernel.raw_lengthscale") for name, _ in m.named_parameters())


def test_lengthscale_priors_enter_the_loss_and_the_gradients():
    """prior_scales / prior_width (projected_lmc.py:135-149, :169-176): Normal / MultivariateNormal priors on the
    lengthscales, initialised at the prior mean; their log-density (counted q times, as gpytorch does) is part of
    the loss and of d loss / d raw_lengthscale."""
    scales = torch.tensor([0.7, 1.3, 0.9, 1.1])
    width = torch.tensor([0.5, 0.25, 0.4, 0.3])
    m = run_case(260, 4, 5, 2, "PLMC_fast", "matern52", decomp=[[0, 2], [1], [3]], prior_scales=scales,
                 prior_width=width, seed=6)
    assert len(list(m.named_priors())) == 3
    sd = m.state_dict()
    assert "covar_module.kernels.0.base_kernel.lengthscale_prior.loc" in sd
    assert "covar_module.kernels.1.base_kernel.lengthscale_prior.scale" in sd
    m2 = run_case(260, 4, 5, 2, "PLMC", "rbf", prior_scales=scales, prior_width=width, outputscales=True, seed=7)
    assert "covar_module.base_kernel.lengthscale_prior._unbroadcasted_scale_tril" in m2.state_dict()


@pytest.mark.parametrize("variant,kernel,os_", [("PLMC", "matern52", False), ("PLMC_fast", "rbf", True)])
def test_inducing_point_model(variant, kernel, os_):
    """ExactGPModel(n_inducing_points=m) (projected_lmc.py:302-303): SGPR loss (low-rank covariance + added trace
    term), gradients to every parameter INCLUDING the inducing points, and the eval-mode prediction with the
    diagonal correction, against the dense oracle."""
    m = run_case(330, 3, 6, 2, variant, kernel, outputscales=os_, n_inducing_points=40, seed=8, grad_tol=1e-6)
    ip = dict(m.named_parameters())["covar_module.inducing_points"]
    assert ip.grad is not None and ip.grad.abs().max().item() > 0
    assert "covar_module.base_kernel.raw_lengthscale" in m.state_dict() or \
        "covar_module.base_kernel.base_kernel.raw_lengthscale" in m.state_dict()


def test_inducing_points_with_an_additive_kernel():
    run_case(280, 4, 5, 2, "PLMC", "matern52", n_inducing_points=150, decomp=[[0, 1], [2, 3]], seed=9, grad_tol=1e-6)


def test_ragged_sizes():
    # n not a multiple of the 128 tile, n < 128, d not a multiple of 4, q = 1, d = 1
    run_case(n=33, d=1, p=4, q=1, variant="PLMC_fast", kernel="matern52", ns=5)
    run_case(n=129, d=5, p=9, q=4, variant="PLMC", kernel="rbf", ns=130)
    run_case(n=384, d=6, p=12, q=5, variant="BDN_diag", kernel="rbf", ns=257)


def test_config1_shape():
    # BASELINE config 1: n=1000, d=6, 50 tasks, 10 latents, RBF, fp64
    run_case(n=1000, d=6, p=50, q=10, variant="PLMC", kernel="rbf", ns=200)
    run_case(n=1000, d=6, p=50, q=10, variant="PLMC_fast", kernel="rbf", ns=64)


def test_sarcos_shape_reduced():
    # config 2 shape (d=21, 7 tasks, 4 latents, Matern-5/2 ARD) at an oracle-sized n
    run_case(n=1500, d=21, p=7, q=4, variant="PLMC", kernel="matern52", ns=100)


def test_api_helpers():
    X, Y, _, _ = synth(100, 3, 6, 2)
    m = make_model(X, Y, 2, variant="PLMC", kernel="rbf").cuda()
    assert m.lscales().shape == (2, 3)
    with pytest.raises(AttributeError):
        m.outputscale()
    assert m.projection_matrix().shape == (6, 2)
    TY = m.project_data(Y.cuda())
    assert TY.shape == (2, 100)
    T = m.projection_matrix()
    assert rel_err(TY, (Y.cuda() @ T).T) < 1e-12
    assert m.projected_noise().shape == (2,)
    assert m.B_tilde().shape == (4, 4)
    assert m.lmc_coefficients().shape == (2, 6)
    with pytest.raises(RuntimeError):
        m(X.cuda()[:50])  # must train on the training inputs


def test_cpu_tensors_fail_loudly():
    from projected_lmc_b200._cabi import PlmcError

    X, Y, _, _ = synth(64, 2, 4, 2)
    m = make_model(X, Y, 2)
    mll = ProjectedLMCmll(m.likelihood, m)
    with pytest.raises(PlmcError):
        mll(m(X), Y)


def test_jitter_retry_matches_oracle():
    # duplicate points + tiny noise -> first factorisation fails, jitter is added to the failing latent only
    X, Y, _, _ = synth(120, 2, 5, 2)
    X[60:] = X[:60]
    m = make_model(X, Y, 2, variant="PLMC_fast", kernel="rbf", perturb=False, noise_thresh=-40.0)
    with torch.no_grad():
        m.likelihood.noise_covar.raw_noise[0] = -60.0
    mc = cpu_copy(m)
    m = m.cuda()
    mll = ProjectedLMCmll(m.likelihood, m)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with gp.settings.cholesky_max_tries(8):
            val = mll(m(X.cuda()), Y.cuda()).item()
        ref = O.mll(oracle_params(mc), X, Y, max_tries=8).item()
    assert m._engine.last_jitter is not None and m._engine.last_jitter[1].item() == 0.0
    assert abs(val - ref) <= 1e-6 * abs(ref)


def test_loo_matches_dense():
    X, Y, _, _ = synth(130, 3, 5, 2)
    m = make_model(X, Y, 2, variant="PLMC", kernel="matern52")
    mc = cpu_copy(m)
    m = m.cuda()
    s2, r = m.compute_loo()
    op = oracle_params(mc)
    with torch.no_grad():
        K = O.gram(op, X, training=False) + torch.diag_embed(O.noise(op)[:, None].expand(-1, 130))
        Kinv = torch.linalg.inv(K)
        TY = O.project_data(op, Y)
        s2_ref = 1.0 / torch.diagonal(Kinv, dim1=1, dim2=2)
        r_ref = (Kinv @ TY.unsqueeze(-1)).squeeze(-1) * s2_ref
    assert rel_err(s2, s2_ref.T) < 1e-8 and rel_err(r, r_ref.T) < 1e-8


def test_kernel_cond_matches_dense():
    X, Y, _, _ = synth(90, 3, 5, 2)
    m = make_model(X, Y, 2, variant="PLMC", kernel="rbf")
    mc = cpu_copy(m)
    m = m.cuda()
    op = oracle_params(mc)
    with torch.no_grad():
        K = O.gram(op, X, training=False) + torch.diag_embed(O.noise(op)[:, None].expand(-1, 90))
        ref = torch.linalg.cond(K)
    assert rel_err(m.kernel_cond(), ref) < 1e-8


def test_fit_loop_matches_oracle_adamw_trajectory():
    """The reference's training loop (AdamW + exponential LR decay, experiments.py:259-284): the CUDA
    path and the CPU oracle must produce the same loss trajectory."""
    from projected_lmc_b200 import fit

    X, Y, _, _ = synth(120, 2, 5, 2, seed=8)
    m = make_model(X, Y, 2, variant="PLMC", kernel="matern52")
    mc = cpu_copy(m)
    m = m.cuda()
    n_iter, lr, lr_min = 12, 1e-2, 1e-3
    out = fit(m, ProjectedLMCmll(m.likelihood, m), X.cuda(), Y.cuda(), n_iter=n_iter, lr=lr, lr_min=lr_min,
              check_every=5, patience=500)
    assert out["n_iter"] == n_iter and out["stopped_at"] is None
    opt = torch.optim.AdamW(mc.parameters(), lr=lr)
    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=float(torch.tensor(lr_min / lr).log().div(n_iter).exp()))
    ref = []
    mc.train()
    for _ in range(n_iter):
        opt.zero_grad()
        loss = -O.mll(oracle_params(mc), X, Y)
        loss.backward()
        opt.step()
        sch.step()
        ref.append(loss.item())
    assert rel_err(out["losses"], torch.tensor(ref)) < 1e-7
    assert out["losses"][-1] < out["losses"][0]


@pytest.mark.parametrize("variant,kernel,n", [("PLMC", "matern52", 300), ("PLMC_fast", "rbf", 1100)])
def test_fit_with_the_whole_step_as_cuda_graphs_follows_the_eager_trajectory(variant, kernel, n):
    """training.fit(cuda_graph=True): forward + backward (QR of the mixing matrix, projection, Gram, Cholesky,
    inverse, sweep, projection terms, autograd) replayed as one CUDA graph, AdamW + LR decay as a second one --
    same loss trajectory and final parameters as the eager loop, and as the oracle's AdamW loop on the CPU."""
    from projected_lmc_b200 import fit

    X, Y, _, _ = synth(n, 3, 6, 2, seed=n)
    n_iter, lr, lr_min = 14, 1e-2, 1e-3
    outs, finals = {}, {}
    for mode in (False, True):
        m = make_model(X, Y, 2, variant=variant, kernel=kernel)
        mc = cpu_copy(m)
        m = m.cuda()
        outs[mode] = fit(m, ProjectedLMCmll(m.likelihood, m), m.train_inputs[0], m.train_y, n_iter=n_iter, lr=lr,
                         lr_min=lr_min, check_every=5, cuda_graph=mode)
        finals[mode] = [p.detach().cpu().clone() for p in m.parameters()]
    assert outs[True]["cuda_graph"], outs[True]["cuda_graph_note"]
    assert not outs[False]["cuda_graph"]
    assert rel_err(outs[True]["losses"], outs[False]["losses"]) < 1e-10
    for a, b in zip(finals[True], finals[False]):
        assert rel_err(a, b) < 1e-7           # capturable AdamW evaluates its bias corrections on the device
    if n <= 300:
        opt = torch.optim.AdamW(mc.parameters(), lr=lr)
        sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=float(torch.tensor(lr_min / lr).log().div(n_iter).exp()))
        ref = []
        for _ in range(n_iter):
            opt.zero_grad()
            loss = -O.mll(oracle_params(mc), X, Y)
            loss.backward()
            opt.step()
            sch.step()
            ref.append(loss.item())
        assert rel_err(outs[True]["losses"], torch.tensor(ref)) < 1e-7


def test_graphed_fit_repeats_a_failed_factorisation_eagerly_with_jitter():
    """Duplicated points and a noise floor of e^-40: the factorisation inside the replayed graph fails, the host
    sees its status between the two graphs and repeats the iteration eagerly (jitter retry) before the optimiser
    graph runs -- same trajectory as the eager loop."""
    from projected_lmc_b200 import fit

    X, Y, _, _ = synth(300, 2, 4, 2, seed=77)
    X[150:] = X[:150]
    outs = {}
    for mode in (False, True):
        m = make_model(X, Y, 2, variant="PLMC", kernel="rbf", perturb=False, noise_thresh=-40.0)
        with torch.no_grad():
            m._base_kernel().raw_lengthscale.fill_(3.0)
            m.likelihood.noise_covar.raw_noise.fill_(-60.0)
        m = m.cuda()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            outs[mode] = fit(m, ProjectedLMCmll(m.likelihood, m), m.train_inputs[0], m.train_y, n_iter=8, lr=1e-3,
                             lr_min=None, check_every=4, cuda_graph=mode)
        assert m._engine.last_jitter is not None and float(m._engine.last_jitter.max()) > 0
    assert outs[True]["cuda_graph"], outs[True]["cuda_graph_note"]
    assert rel_err(outs[True]["losses"], outs[False]["losses"]) < 1e-8


def test_fit_falls_back_to_eager_launches_when_the_step_cannot_be_captured():
    """A step with a host read in it cannot be captured: fit() warns, continues eagerly with the same trajectory,
    and torch's CUDA generator is usable afterwards (a capture that ends in an error leaves it flagged)."""
    from projected_lmc_b200 import fit

    X, Y, _, _ = synth(200, 2, 4, 2, seed=3)
    outs = {}
    for leaky in (False, True):
        m = make_model(X, Y, 2, variant="PLMC", kernel="rbf").cuda()
        mll = ProjectedLMCmll(m.likelihood, m)
        if leaky:
            orig = mll.forward

            def fwd(*a, _orig=orig, **k):
                out = _orig(*a, **k)
                float(out.detach().cpu())                  # host read: not capturable
                return out
            mll.forward = fwd
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            outs[leaky] = fit(m, mll, m.train_inputs[0], m.train_y, n_iter=8, lr=1e-3, lr_min=None, cuda_graph=True)
        if leaky:
            assert not outs[leaky]["cuda_graph"] and any("could not be captured" in str(x.message) for x in w)
    assert outs[False]["cuda_graph"]
    assert rel_err(outs[True]["losses"], outs[False]["losses"]) < 1e-10
    assert torch.isfinite(torch.randn(4, device="cuda")).all()


@pytest.mark.parametrize("variant", ["PLMC", "PLMC_fast"])
def test_one_shared_qr_per_loss_evaluation_gives_the_gradients_of_two(variant):
    """ProjectedLMCmll.forward factorises H once for project_data and the projection terms (qr_once); the reference
    factorises twice.  Same loss bits, gradients equal to rounding."""
    import contextlib

    X, Y, _, _ = synth(400, 3, 6, 2, seed=21)
    m = make_model(X, Y, 2, variant=variant, kernel="matern52").cuda()
    Xg, Yg = X.cuda(), Y.cuda()
    res = []
    for shared in (True, False):
        if not shared:
            m.lmc_coefficients.qr_once = contextlib.nullcontext          # two factorisations, as in the reference
        for p in m.parameters():
            p.grad = None
        loss = -ProjectedLMCmll(m.likelihood, m)(m(Xg), Yg)
        loss.backward()
        res.append((loss.item(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    assert abs(res[0][0] - res[1][0]) <= 1e-14 * abs(res[1][0])
    for k in res[0][1]:
        assert rel_err(res[0][1][k], res[1][1][k]) < 1e-12, k


def test_fit_plateau_stop_rule():
    from projected_lmc_b200 import fit

    X, Y, _, _ = synth(64, 2, 4, 2, seed=2)
    m = make_model(X, Y, 2, variant="PLMC_fast", kernel="rbf").cuda()
    out = fit(m, ProjectedLMCmll(m.likelihood, m), X.cuda(), Y.cuda(), n_iter=60, lr=1e-6, lr_min=None,
              loss_thresh=1e-2, patience=7, check_every=4)
    # with a vanishing learning rate every step is a plateau step: the rule fires at iteration patience + 1
    assert out["stopped_at"] == 8 and out["n_iter"] <= 12


def test_eval_cache_is_invalidated_by_parameter_updates():
    X, Y, Xs, _ = synth(100, 2, 5, 2, ns=20)
    m = make_model(X, Y, 2, variant="PLMC", kernel="rbf").cuda()
    m.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = m(Xs.cuda()).mean.clone()
        b = m(Xs.cuda()).mean.clone()
        assert torch.equal(a, b)
        key = m._pred_key
        m.covar_module.raw_lengthscale.add_(0.3)
        c = m(Xs.cuda()).mean
        assert m._pred_key != key and not torch.allclose(a, c)


def test_float32_small_model_within_fp32_tolerance():
    X, Y, Xs, _ = synth(80, 2, 4, 2, ns=10)
    m = make_model(X, Y, 2, variant="PLMC_fast", kernel="rbf")
    ref = -O.mll(oracle_params(cpu_copy(m)), X, Y)
    m32 = m.float().cuda()
    loss = -ProjectedLMCmll(m32.likelihood, m32)(m32(X.float().cuda()), Y.float().cuda())
    assert loss.dtype == torch.float32
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())       # north_star fp32 tolerance
    loss.backward()
    assert all(p.grad is not None and p.grad.dtype == torch.float32 for p in m32.parameters())


def test_float32_model_runs_the_fp32_grade_on_the_tensor_path_within_1e_minus_4():
    """dtype = float32 (the reference's GPU default, experiments.py:4-8): 32-bit operands on the INT8 path
    (10 moduli / 4 digit planes), FP64 storage; north_star tolerance 1e-4 for MLL, gradients and predictions."""
    X, Y, Xs, _ = synth(1500, 4, 6, 3, seed=12, ns=40)
    m = make_model(X, Y, 3, variant="PLMC", kernel="matern52")
    mc = cpu_copy(m)
    ref = -O.mll(oracle_params(mc), X, Y)
    ref.backward()
    m32 = m.float().cuda()
    assert m32._engine.grade == "fp32"
    loss = -ProjectedLMCmll(m32.likelihood, m32)(m32(X.float().cuda()), Y.float().cuda())
    loss.backward()
    cfg = m32._engine.cfg_main
    assert cfg is not None and cfg.precision == m32._engine.rns_moduli_f32 and m32._engine.cfg_kinv.precision == cfg.precision
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    refg = dict(mc.named_parameters())
    for name, prm in m32.named_parameters():
        if refg[name].grad is not None:
            assert rel_err(prm.grad, refg[name].grad) <= 1e-4, name
    m32.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pred = m32.full_likelihood()(m32(Xs.float().cuda()))
        mean_ref, _, var_ref = O.predict(oracle_params(mc), X, Y, Xs)
    assert pred.mean.dtype == torch.float32
    assert rel_err(pred.mean, mean_ref) <= 1e-4 and rel_err(pred.variance, var_ref) <= 1e-4
    assert m32.double()._engine.grade == "fp64"


def test_more_than_44_dimensions_in_one_kernel_is_refused_with_the_remedy():
    from projected_lmc_b200 import PlmcError

    X, Y, _, _ = synth(200, 50, 4, 2, seed=3)
    m = make_model(X, Y, 2, variant="PLMC_fast", kernel="rbf").cuda()
    with pytest.raises(PlmcError, match="decomp"):
        ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
    # ... and the remedy works: two additive groups of 25 dimensions
    m2 = make_model(X, Y, 2, variant="PLMC_fast", kernel="rbf", decomp=[list(range(25)), list(range(25, 50))]).cuda()
    loss = -ProjectedLMCmll(m2.likelihood, m2)(m2(X.cuda()), Y.cuda())
    loss.backward()
    assert torch.isfinite(loss)


@pytest.mark.parametrize("n,variant", [(700, "PLMC"), (1500, "PLMC_fast")])
def test_cuda_graph_replay_of_the_engine_step_is_bit_identical_to_eager(n, variant):
    """Small problems replay the engine part of the step as two CUDA graphs (LatentEngine.graph_max_order): same
    bits as the eager launches over several optimiser steps (inputs are re-read on every replay), and the jitter
    retry still works from inside the graphed path."""
    from projected_lmc_b200.engine import LatentEngine

    X, Y, _, _ = synth(n, 3, 5, 2, seed=n)
    Xg, Yg = X.cuda(), Y.cuda()
    hist = {}
    old = LatentEngine.graph_max_order
    try:
        for order in (0, 4096):
            LatentEngine.graph_max_order = order
            m = make_model(X, Y, 2, variant=variant, kernel="matern52").cuda()
            mll = ProjectedLMCmll(m.likelihood, m)
            opt = torch.optim.SGD(m.parameters(), lr=1e-3)
            out = []
            for it in range(5):
                opt.zero_grad(set_to_none=True)
                loss = -mll(m(Xg), Yg)
                loss.backward()
                out.append((loss.item(), [p.grad.clone() for p in m.parameters()]))
                opt.step()
            hist[order] = out
            if order:
                assert any(isinstance(v, dict) for v in m._engine._graphs.values())      # really captured
    finally:
        LatentEngine.graph_max_order = old
    for (l0, g0), (l1, g1) in zip(hist[0], hist[4096]):
        assert l0 == l1 and all(torch.equal(a, b) for a, b in zip(g0, g1))


def test_jitter_retry_from_the_graphed_path():
    X, Y, _, _ = synth(300, 2, 4, 2, seed=77)
    X[150:] = X[:150]
    m = make_model(X, Y, 2, variant="PLMC", kernel="rbf", perturb=False, noise_thresh=-40.0)
    with torch.no_grad():
        m._base_kernel().raw_lengthscale.fill_(3.0)
        m.likelihood.noise_covar.raw_noise.fill_(-60.0)
    mc = cpu_copy(m)
    m = m.cuda()
    losses = []
    with gp.settings.cholesky_max_tries(8), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for it in range(3):                      # eager warm-up, capture, replay: all three must retry and agree
            for p in m.parameters():
                p.grad = None
            loss = -ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
            loss.backward()
            losses.append(loss.item())
            assert m._engine.last_jitter is not None and float(m._engine.last_jitter.max()) > 0
        ref = -O.mll(oracle_params(mc), X, Y, max_tries=8)
    assert losses[0] == losses[1] == losses[2]
    assert abs(losses[0] - ref.item()) <= 1e-5 * abs(ref.item())


def test_not_psd_error_after_max_tries():
    from projected_lmc_b200 import NotPSDError

    X, Y, _, _ = synth(100, 2, 4, 2)
    X[50:] = X[:50]                       # exactly duplicated points, no noise to speak of -> singular
    m = make_model(X, Y, 2, variant="PLMC_fast", kernel="rbf", perturb=False, noise_thresh=-60.0)
    with torch.no_grad():
        m.likelihood.noise_covar.raw_noise.fill_(-80.0)
    m = m.cuda()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # a jitter far below the rounding level cannot repair the singular matrix
        with gp.settings.cholesky_max_tries(2), gp.settings.cholesky_jitter(1e-40), pytest.raises(NotPSDError):
            ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())


def test_nan_inputs_raise_like_psd_safe_cholesky():
    """linear_operator's psd_safe_cholesky raises NanError on a matrix with NaN entries before factoring."""
    from projected_lmc_b200 import NanError

    X, Y, _, _ = synth(300, 3, 4, 2)
    X[17, 1] = float("nan")
    m = make_model(X, Y, 2, variant="PLMC", kernel="rbf").cuda()
    with pytest.raises(NanError):
        ProjectedLMCmll(m.likelihood, m)(m(m.train_inputs[0]), Y.cuda())


def test_non_finite_test_points_give_nan_rows_only():
    X, Y, Xs, _ = synth(400, 3, 4, 2, ns=20)
    m = make_model(X, Y, 2, variant="PLMC", kernel="matern52").cuda()
    m.eval()
    with torch.no_grad():
        good = m(Xs.cuda())
        Xb = Xs.clone()
        Xb[3, 0] = float("nan")
        Xb[11, 2] = float("inf")
        pred = m(Xb.cuda())
    keep = torch.ones(20, dtype=torch.bool)
    keep[[3, 11]] = False
    assert torch.isnan(pred.mean[~keep]).all() and torch.isnan(pred.variance[~keep]).all()
    assert torch.equal(pred.mean[keep], good.mean[keep]) and torch.equal(pred.variance[keep], good.variance[keep])
