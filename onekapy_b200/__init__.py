"""onekapy_b200 -- B200-native Monte Carlo capture-zone hot path of OnekaPy.

  onekapy_b200.csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/oneka_b200.h)
  onekapy_b200._cabi      ctypes binding of liboneka_b200.so
  onekapy_b200.engine     torch plumbing: device memory, stream, lattice choice, cropping
  onekapy_b200.parallel   realization sharding + the one NCCL allreduce of the count grid
  onekapy_b200.host.*     mirror of the reference's Python interface (also importable as `oneka.*`)
"""
from ._cabi import OnekaError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
