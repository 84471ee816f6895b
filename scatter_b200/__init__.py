"""scatter_b200 -- B200-native hot path of SCATTER (assembly + time integration) behind the reference's own API."""
from .scatter import scatter, Solver  # noqa: F401

__version__ = "0.1.0"
