"""Drop-in module `oneka.utilities` (same public names as the reference's oneka/utilities.py);
the implementation lives in onekapy_b200.host.utilities."""
from onekapy_b200.host.utilities import *  # noqa: F401,F403
from onekapy_b200.host import utilities as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
