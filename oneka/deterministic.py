"""Drop-in module `oneka.deterministic` (same public names as the reference's oneka/deterministic.py);
the implementation lives in onekapy_b200.host.deterministic."""
from onekapy_b200.host.deterministic import *  # noqa: F401,F403
from onekapy_b200.host import deterministic as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
