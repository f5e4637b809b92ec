"""Lengthscale priors of handle_covar_ (projected_lmc.py:135-149): gpytorch.priors.NormalPrior and
MultivariateNormalPrior restricted to what the reference builds (a diagonal covariance).  Buffers carry
gpytorch's names so ``state_dict()`` keys match; the densities are plain torch on a [q, 1, d] tensor
(host-side O(q d) work: autograd reaches the raw lengthscales)."""
from __future__ import annotations

import math

import torch


class Prior(torch.nn.Module):
    pass


class NormalPrior(Prior):
    def __init__(self, loc, scale):
        super().__init__()
        self.register_buffer("loc", torch.as_tensor(loc).clone())
        self.register_buffer("scale", torch.as_tensor(scale).clone())

    @property
    def mean(self):
        return self.loc

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """Elementwise log N(x; loc, scale^2) (torch.distributions.Normal.log_prob)."""
        var = self.scale ** 2
        return -((x - self.loc) ** 2) / (2 * var) - torch.log(self.scale) - 0.5 * math.log(2 * math.pi)


class MultivariateNormalPrior(Prior):
    def __init__(self, loc, covariance_matrix):
        super().__init__()
        loc = torch.as_tensor(loc).clone()
        cov = torch.as_tensor(covariance_matrix).clone()
        self.register_buffer("loc", loc)
        self.register_buffer("_unbroadcasted_scale_tril", torch.linalg.cholesky(cov))

    @property
    def mean(self):
        return self.loc

    @property
    def covariance_matrix(self):
        L = self._unbroadcasted_scale_tril
        return L @ L.transpose(-1, -2)

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """log N(x; loc, Sigma) over the last dimension (torch.distributions.MultivariateNormal.log_prob)."""
        L = self._unbroadcasted_scale_tril
        diff = (x - self.loc).unsqueeze(-1)
        z = torch.linalg.solve_triangular(L.expand(*diff.shape[:-2], *L.shape), diff, upper=False).squeeze(-1)
        half_logdet = torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
        d = x.shape[-1]
        return -0.5 * (z ** 2).sum(-1) - half_logdet - 0.5 * d * math.log(2 * math.pi)
