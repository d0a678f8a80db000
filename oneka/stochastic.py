"""Drop-in module `oneka.stochastic` (same public names as the reference's oneka/stochastic.py);
the implementation lives in onekapy_b200.host.stochastic."""
from onekapy_b200.host.stochastic import *  # noqa: F401,F403
from onekapy_b200.host import stochastic as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
