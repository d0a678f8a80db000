"""Drop-in module `oneka.probabilityfield` (same public names as the reference's oneka/probabilityfield.py);
the implementation lives in onekapy_b200.host.probabilityfield."""
from onekapy_b200.host.probabilityfield import *  # noqa: F401,F403
from onekapy_b200.host import probabilityfield as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
