"""Mixing-matrix module and parametrisations of the projected model (host side, tiny).

Mirrors LMCMixingMatrix (projected_lmc.py:819-890) and the four parametrisation
classes (:207-258): same names, shapes, parameter names and error behaviour.  The QR
stays in torch so autograd delivers dH."""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor


def _diag_idx(mat: Tensor):
    k = mat.shape[-1]
    return torch.arange(k, device=mat.device)


class ScalarParam(torch.nn.Module):
    """Every entry equals the (clamped) mean of the raw vector."""

    def __init__(self, bounds: Sequence[float] = (1e-16, 1e16)):
        super().__init__()
        self.bounds = bounds

    def forward(self, X: Tensor) -> Tensor:
        lo, hi = self.bounds
        return X.mean().clamp(lo, hi) * torch.ones_like(X)

    def right_inverse(self, A: Tensor) -> Tensor:
        return A


class PositiveDiagonalParam(torch.nn.Module):
    """Diagonal matrix with exp() of the raw diagonal."""

    def forward(self, X: Tensor) -> Tensor:
        return torch.diag_embed(torch.diag(X).exp())

    def right_inverse(self, A: Tensor) -> Tensor:
        return torch.diag_embed(torch.diag(A).log())


class UpperTriangularParam(torch.nn.Module):
    """Upper-triangular matrix whose diagonal is exp() of the raw diagonal."""

    def forward(self, X: Tensor) -> Tensor:
        out = X.triu()
        i = _diag_idx(out)
        out[i, i] = out[i, i].exp()
        return out

    def right_inverse(self, A: Tensor) -> Tensor:
        i = _diag_idx(A)
        A[i, i] = A[i, i].log()
        return A


class LowerTriangularParam(torch.nn.Module):
    """Cholesky-factor parametrisation: lower-triangular, diagonal exp(clamp(raw))."""

    def __init__(self, bounds: Sequence[float] = (1e-16, 1e16)):
        super().__init__()
        self.bounds = bounds

    def forward(self, X: Tensor) -> Tensor:
        out = X.tril()
        i = _diag_idx(out)
        lo, hi = self.bounds
        out[i, i] = out[i, i].clamp(lo, hi).exp()
        return out

    def right_inverse(self, A: Tensor) -> Tensor:
        i = _diag_idx(A)
        A[i, i] = A[i, i].log()
        return A


class LMCMixingMatrix(torch.nn.Module):
    """Parametrised mixing matrix H = Q R.

    bulk=True stores the raw product H (padded to p x p in mode 'Q_plus') and
    re-factorises it on every call; bulk=False stores Q_plus and R separately so
    that torch parametrisations (orthogonal / triangular) can be attached."""

    def __init__(self, Q_plus: Tensor, R: Tensor, bulk: bool = True):
        super().__init__()
        p, cols = Q_plus.shape
        q = R.shape[0]
        if cols == p:
            self.mode = "Q_plus"
        elif cols == q:
            self.mode = "Q"
        else:
            raise ValueError("Wrong dimensions for Q_plus : should be n_tasks x n_tasks or n_tasks x n_latents")
        self.n_latents, self.n_tasks = q, p
        self._size = torch.Size([q, p])
        self.bulk = bulk
        if bulk:
            if self.mode == "Q_plus":
                R_full = torch.eye(p)  # default dtype / CPU on purpose (reference quirk, :845)
                R_full[:q, :q] = R
                H = Q_plus @ R_full
            else:
                H = Q_plus @ R
            self.register_parameter("H", torch.nn.Parameter(H, requires_grad=True))
        else:
            self.register_parameter("Q_plus", torch.nn.Parameter(Q_plus, requires_grad=True))
            self.register_parameter("R", torch.nn.Parameter(R, requires_grad=True))

    def Q(self) -> Tensor:
        return self.Q_plus[:, : self.n_latents] if self.mode == "Q_plus" else self.Q_plus

    def Q_orth(self) -> Tensor:
        return self.Q_plus[:, self.n_latents:]

    def QR(self):
        """(Q [p,q], R [q,q], Q_orth [p,p-q] | None)."""
        if not self.bulk:
            return self.Q(), self.R, self.Q_orth()
        memo = self.__dict__.get("_qr_memo")
        if memo is not None and memo[0] is not None:
            return memo[0]
        q = self.n_latents
        Qf, Rf = torch.linalg.qr(self.H)
        out = (Qf[:, :q], Rf[:q, :q], Qf[:, q:]) if self.mode == "Q_plus" else (Qf, Rf, None)
        if memo is not None:
            memo[0] = out
        return out

    def qr_once(self):
        """Context manager: inside it the factorisation of H is computed once and shared by every caller.  The
        reference factorises H twice per loss evaluation (project_data, :1015, and the projection terms, :1208); the
        two results are identical, so ProjectedLMCmll.forward evaluates both under one factorisation (one QR and one
        QR-backward less per iteration -- 8 % of the GPU time of a launch-bound step).  The memo lives only for the
        duration of the block: it never outlives the autograd graph it belongs to."""
        module = self

        class _Scope:
            def __enter__(self_inner):
                module.__dict__["_qr_memo"] = [None]

            def __exit__(self_inner, *exc):
                module.__dict__["_qr_memo"] = None
                return False

        return _Scope()

    def forward(self) -> Tensor:
        """H^T, shape n_latents x n_tasks."""
        if self.bulk:
            return self.H.T if self.mode == "Q" else self.H[:, : self.n_latents].T
        return (self.Q() @ self.R).T

    def size(self, int=None):
        return self._size[int] if int else self._size
