"""``from projectedlmc.projected_lmc import ProjectedGPModel`` / ``from projected_lmc import *``
(the flat import used by experiments.py:1) resolve to the B200-native implementation."""
from projected_lmc_b200 import *  # noqa: F401,F403
from projected_lmc_b200 import gp  # noqa: F401
