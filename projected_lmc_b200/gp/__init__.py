"""``gp``: the slice of the gpytorch namespace the projected-LMC path touches.

Usage mirrors the reference (``import gpytorch as gp``):
``from projected_lmc_b200 import gp; gp.kernels.MaternKernel``."""
from . import constraints, distributions, kernels, likelihoods, means, mlls, priors, settings  # noqa: F401
