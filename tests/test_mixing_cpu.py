"""CPU: LMCMixingMatrix.qr_once -- one factorisation of H shared inside the scope, none kept outside it."""
import torch

from projected_lmc_b200.mixing import LMCMixingMatrix


def _module(p=6, q=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    Qp, _ = torch.linalg.qr(torch.randn(p, p, generator=g, dtype=torch.float64))
    R = torch.triu(torch.randn(q, q, generator=g, dtype=torch.float64)) + 2 * torch.eye(q, dtype=torch.float64)
    torch.set_default_dtype(torch.float64)
    return LMCMixingMatrix(Qp, R, bulk=True)


def test_qr_is_shared_inside_the_scope_and_recomputed_outside(monkeypatch):
    m = _module()
    calls = []
    real = torch.linalg.qr

    def counting(*a, **k):
        calls.append(1)
        return real(*a, **k)

    monkeypatch.setattr(torch.linalg, "qr", counting)
    m.QR(); m.QR()
    assert len(calls) == 2                       # no memo outside a scope
    with m.qr_once():
        a = m.QR()
        b = m.QR()
        assert len(calls) == 3 and all(x is y for x, y in zip(a, b))
    m.QR()
    assert len(calls) == 4                       # the memo died with the scope
    with m.qr_once():
        c = m.QR()
    assert len(calls) == 5 and c[0] is not a[0]


def test_gradients_through_one_shared_factorisation_equal_those_through_two():
    m = _module(seed=3)

    def loss(shared):
        m.H.grad = None

        def body():
            Q, R, Qo = m.QR()
            f = (Q @ R).pow(2).sum() + torch.linalg.solve_triangular(R.T, Q, upper=False, left=False).sum()
            Q2, R2, Qo2 = m.QR()
            return f - 0.5 * torch.log(torch.diagonal(R2) ** 2).sum() + (Qo2.T @ Qo2).diagonal().sum() + Qo2.sum()

        if shared:
            with m.qr_once():
                out = body()
        else:
            out = body()
        out.backward()
        return out.item(), m.H.grad.clone()

    l1, g1 = loss(True)
    l2, g2 = loss(False)
    assert abs(l1 - l2) <= 1e-14 * abs(l2)
    assert (g1 - g2).abs().max() <= 1e-12 * g2.abs().max()
