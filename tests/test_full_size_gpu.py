"""GPU, BASELINE.json full size (configs[1]: n=44,484, d=21, 7 tasks, 4 latents, Matern-5/2, fp64).

No oracle can run at this size, so parity is checked through size-independent properties:
the analytic gradient of one training step must agree with a central finite difference of the
loss along a random direction in the raw-parameter space, and a repeated step must be
bit-identical (fixed-order reductions everywhere)."""
import pytest
import torch

from projected_lmc_b200 import ProjectedLMCmll

from .helpers import make_model

pytestmark = pytest.mark.gpu


def test_c2_full_size_directional_derivative_and_determinism():
    if torch.cuda.get_device_properties(0).total_memory < 90e9:
        pytest.skip("needs ~70 GB of HBM")
    import bench

    n, d, p, q = 44484, 21, 7, 4
    X, Y = bench.make_data(n, d, p, q, seed=0)
    m = make_model(X, Y, q, variant="PLMC", kernel="matern52").cuda()
    Xg, Yg = m.train_inputs[0], m.train_y
    mll = ProjectedLMCmll(m.likelihood, m)
    params = [prm for prm in m.parameters() if prm.requires_grad]

    def loss_only():
        with torch.no_grad():
            return float(-mll(m(Xg), Yg))

    loss = -mll(m(Xg), Yg)
    loss.backward()
    g1 = [prm.grad.clone() for prm in params]
    for prm in params:
        prm.grad = None
    loss2 = -mll(m(Xg), Yg)
    loss2.backward()
    assert loss.item() == loss2.item()
    assert all(torch.equal(a, prm.grad) for a, prm in zip(g1, params))          # deterministic

    gen = torch.Generator().manual_seed(3)
    dirs = [torch.randn(prm.shape, generator=gen, dtype=torch.float64).to(prm.device) for prm in params]
    analytic = sum(float((g * v).sum()) for g, v in zip(g1, dirs))
    eps = 1e-4
    with torch.no_grad():
        for prm, v in zip(params, dirs):
            prm.add_(eps * v)
        lp = loss_only()
        for prm, v in zip(params, dirs):
            prm.add_(-2 * eps * v)
        lm = loss_only()
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - analytic) <= 1e-5 * max(1.0, abs(analytic)), (fd, analytic)
