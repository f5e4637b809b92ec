"""Base class mirroring gpytorch.mlls.ExactMarginalLogLikelihood's constructor."""
import torch

from .likelihoods import _GaussianLikelihoodBase


class MarginalLogLikelihood(torch.nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model


class ExactMarginalLogLikelihood(MarginalLogLikelihood):
    def __init__(self, likelihood, model):
        if not isinstance(likelihood, _GaussianLikelihoodBase):
            raise RuntimeError("Likelihood must be Gaussian for exact inference")
        super().__init__(likelihood, model)

    def _add_other_terms(self, res, params):
        # added-loss terms exist only for inducing-point kernels and log-priors only
        # with prior_scales (projected_lmc.py:135-149); both are out of scope here.
        return res
