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
        """gpytorch 1.11 ExactMarginalLogLikelihood._add_other_terms: added-loss terms of the model (only the
        inducing-point kernel has one) and the log-density of every registered prior (only the lengthscale priors of
        handle_covar_, projected_lmc.py:135-149).  As in gpytorch, each SUMMED prior term is added to the whole
        [q]-shaped result -- every latent's entry receives it -- so after the caller's ``.sum()`` it counts q times;
        the reference inherits that, and so does this."""
        model = self.model
        if hasattr(model, "added_loss_terms"):
            for term in model.added_loss_terms():
                res = res + term.loss(*params)
        if hasattr(model, "named_priors"):
            for _, module, prior, closure in model.named_priors():
                res = res + prior.log_prob(closure(module)).sum()
        return res
