"""Minimal distribution objects returned by the model (no gpytorch dependency)."""
from __future__ import annotations

import torch

from . import settings


class Distribution:
    pass


class MultivariateNormal(Distribution):
    """Train-mode handle returned by ``model(X)``: the batched latent prior N(0, K_l).

    Like gpytorch's lazy ``MultivariateNormal(mean, LazyEvaluatedKernelTensor)`` it
    computes nothing; the MLL hands it to the CUDA engine (projected_lmc.py:1130-1131)."""

    def __init__(self, model, x, noise_likelihood=None):
        self.model = model
        self.x = x
        self.noise_likelihood = noise_likelihood

    def with_noise(self, likelihood):
        return MultivariateNormal(self.model, self.x, likelihood)

    @property
    def batch_shape(self):
        return torch.Size([self.model.n_latents])

    @property
    def event_shape(self):
        return torch.Size([self.x.shape[-2]])

    @property
    def mean(self):
        return torch.zeros(self.model.n_latents, self.x.shape[-2], dtype=self.x.dtype, device=self.x.device)

    def log_prob(self, value):
        if self.noise_likelihood is None:
            raise RuntimeError("log_prob of the noise-free latent prior is not defined; call likelihood(dist) first")
        return self.model._latent_log_prob(value)


class MultitaskMultivariateNormal(Distribution):
    """Eval-mode result: task means [n*, p] and marginal variances [n*, p].

    The reference builds the full (n* p) x (n* p) covariance as a dense Kronecker sum
    (projected_lmc.py:1149-1155); every consumer only reads its diagonal, which is what
    is stored here."""

    def __init__(self, mean, variance):
        self._mean = mean
        self._var = variance

    @property
    def mean(self):
        return self._mean

    @property
    def variance(self):
        return self._var.clamp_min(settings.min_variance.value())

    @property
    def stddev(self):
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = self.stddev * 2
        return self.mean - s2, self.mean + s2

    def add_task_noise(self, task_var):
        return MultitaskMultivariateNormal(self._mean, self._var + task_var[None, :])

    @property
    def covariance_matrix(self):
        raise NotImplementedError("only marginal variances are computed on the B200 path (SURVEY.md 8a row a7)")
