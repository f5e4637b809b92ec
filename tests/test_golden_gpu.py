"""GPU: the CUDA path (through the model API / C ABI) against the committed reference runs
(tests/golden/reference_runs.pt: the reference's own ProjectedGPModel / ProjectedLMCmll source
executed in the build container, see tests/golden/make_golden.py)."""
import os
import warnings

import pytest
import torch

from projected_lmc_b200 import ProjectedLMCmll, gp

from .helpers import rel_err
from .test_golden_cpu import GOLD, build_from_case, case_keys

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def runs():
    return torch.load(os.path.join(GOLD, "reference_runs.pt"), weights_only=False)


@pytest.mark.parametrize("key", case_keys())
def test_cuda_path_vs_reference_run(runs, key):
    m, case = build_from_case(runs, key)
    m = m.cuda()
    X, Y, Xs = runs["X"].cuda(), runs["Y"].cuda(), runs["Xs"].cuda()
    m.train()
    mll = ProjectedLMCmll(m.likelihood, m)
    with gp.settings.cholesky_max_tries(8):
        loss = -mll(m(X), Y)
    loss.backward()
    assert abs(loss.item() - case["loss"]) <= 1e-9 * abs(case["loss"])          # north_star: 1e-8
    for got, want in zip(mll.proj_term_list, case["proj_terms"]):
        assert abs(float(got) - want) <= 1e-9 * max(1.0, abs(want))
    for name, prm in m.named_parameters():
        want = case["grads"][name]
        if want is None:
            assert prm.grad is None or prm.grad.abs().max().item() == 0.0
        else:
            assert rel_err(prm.grad, want) <= 1e-7, name                        # north_star: 1e-6
    assert rel_err(m.project_data(Y), case["TY"]) <= 1e-11
    m.eval()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lat = m(Xs)
        obs = m.full_likelihood()(lat)
    assert rel_err(lat.mean, case["mean"]) <= 1e-7
    assert rel_err(lat.variance, case["var_f"]) <= 1e-7
    assert rel_err(obs.variance, case["var_y"]) <= 1e-7
