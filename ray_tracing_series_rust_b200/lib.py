"""Build and load librtb200.so (in-tree, so the built file travels to the GPU box)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("RTB200_LIB") or os.path.join(_HERE, "librtb200.so")  # RTB200_LIB: A/B another build of the same library

_api = None


def build(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile); cross-compiles without a GPU."""
    p = subprocess.run(["make", "-C", CSRC, "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or p.returncode != 0:
        print(p.stdout)
    if p.returncode != 0:
        raise RuntimeError("building librtb200.so failed:\n" + p.stdout[-4000:])
    return LIB_PATH


def load() -> capi.Api:
    """Open librtb200.so and bind every symbol include/rtb200.h declares.  Raises if it is missing."""
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        _api = capi.Api(lib, "rt_", extra=capi.DEVICE_SIGNATURES)
    return _api


def new_scene() -> capi.Scene:
    return capi.Scene(load())
