"""Drop-in import name of the reference package (``projectedlmc/__init__.py:1`` re-exports
everything): ``from projectedlmc import *`` gives the B200-native classes."""
from projected_lmc_b200 import *  # noqa: F401,F403
from projected_lmc_b200 import gp  # noqa: F401
