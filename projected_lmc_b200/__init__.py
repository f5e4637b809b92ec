"""projected_lmc_b200: B200-native (sm_100a) engine for the projected-LMC hot path.

The public model API (``ProjectedGPModel``, ``ProjectedLMCmll`` ...) mirrors
``projectedlmc/projected_lmc.py`` of the reference; the numerical work is done by
hand-written CUDA kernels behind the C ABI in ``include/plmc_b200.h``.
"""
from . import gp  # noqa: F401
from ._cabi import PlmcError  # noqa: F401
from .engine import LatentEngine, NanError, NotPSDError  # noqa: F401
from .mixing import (LMCMixingMatrix, LowerTriangularParam, PositiveDiagonalParam, ScalarParam,  # noqa: F401
                     UpperTriangularParam)
from .mll import ProjectedLMCmll, projection_terms  # noqa: F401
from .models import ExactGPModel, ProjectedGPModel, handle_covar_, init_lmc_coefficients  # noqa: F401
from .training import fit  # noqa: F401

__version__ = "0.1.0"
__all__ = [
    "gp", "ProjectedGPModel", "ProjectedLMCmll", "ExactGPModel", "LMCMixingMatrix", "ScalarParam",
    "PositiveDiagonalParam", "UpperTriangularParam", "LowerTriangularParam", "handle_covar_",
    "init_lmc_coefficients", "LatentEngine", "NotPSDError", "NanError", "PlmcError", "projection_terms", "fit",
]
