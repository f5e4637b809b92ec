"""projected_lmc_b200: B200-native (sm_100a) engine for the projected-LMC hot path.

The public model API (``ProjectedGPModel``, ``ProjectedLMCmll`` ...) mirrors
``projectedlmc/projected_lmc.py`` of the reference; the numerical work is done by
hand-written CUDA kernels behind the C ABI in ``include/plmc_b200.h``.
"""
__version__ = "0.1.0"
