"""Drop-in module `oneka.capturezone` (same public names as the reference's oneka/capturezone.py);
the implementation lives in onekapy_b200.host.capturezone."""
from onekapy_b200.host.capturezone import *  # noqa: F401,F403
from onekapy_b200.host import capturezone as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
