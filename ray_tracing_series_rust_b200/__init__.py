"""rtb200 — B200-native path-tracing hot path behind the ray-tracing-series-rust scene API.

The product is the C-ABI shared library `librtb200.so` (include/rtb200.h) built from csrc/
(host C++ flatten/BVH/IO + hand-written sm_100a CUDA).  This package is the thin Python host side:
`capi` binds the C-ABI with ctypes, `lib.load()` opens the in-tree library and FAILS LOUDLY if it
has not been built — there is no CPU or PyTorch fallback anywhere in the product path.
"""
from . import capi  # noqa: F401
from .lib import LIB_PATH, build, load, new_scene  # noqa: F401

__all__ = ["capi", "LIB_PATH", "build", "load", "new_scene"]
