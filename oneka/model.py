"""Drop-in module `oneka.model` (same public names as the reference's oneka/model.py);
the implementation lives in onekapy_b200.host.model."""
from onekapy_b200.host.model import *  # noqa: F401,F403
from onekapy_b200.host import model as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
