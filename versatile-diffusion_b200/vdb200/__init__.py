"""vdb200 — Python binding of libvdb200.so (the sm_100a kernels of the VD sampling hot path).

The library is loaded eagerly and loudly: there is no CPU or library fallback (north-star rule).
"""
from ._lib import lib, check, VdbError, LIB_PATH  # noqa: F401
from . import ops  # noqa: F401
