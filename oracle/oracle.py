"""ctypes binding of the CPU oracle (oracle/liboracle_rt.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ray_tracing_series_rust_b200 import capi  # noqa: E402  (binding table only; loads no library)

LIB_PATH = os.path.join(_HERE, "liboracle_rt.so")

c_d3 = capi.c_d3
_P = C.c_void_p

KAT_SIGNATURES = {
    "kat_philox": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "kat_vec3": (None, [C.c_int32, c_d3, c_d3, C.c_double, c_d3]),
    "kat_vec3_scalar": (C.c_double, [C.c_int32, c_d3, c_d3]),
    "kat_aabb_hit": (C.c_int32, [c_d3, c_d3, c_d3, c_d3, C.c_double, C.c_double]),
    "kat_reflectance": (C.c_double, [C.c_double, C.c_double]),
    "kat_refract": (None, [c_d3, c_d3, C.c_double, c_d3]),
    "kat_reflect": (None, [c_d3, c_d3, c_d3]),
    "kat_normalized_color": (None, [c_d3, C.c_uint32, c_d3]),
    "kat_moving_center": (None, [c_d3, c_d3, C.c_double, C.c_double, C.c_double, c_d3]),
    "kat_camera": (C.c_int32, [_P, c_d3]),
    "kat_camera_ray": (C.c_int32, [_P, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "kat_texture_value": (C.c_int32, [_P, C.c_int32, C.c_double, C.c_double, c_d3, c_d3]),
    "kat_perlin": (C.c_int32, [_P, C.c_int32, c_d3, c_d3, c_d3]),
    "kat_gravity_table": (C.c_int32, [C.c_double, C.c_double, C.c_double, C.c_int32, c_d3, C.POINTER(C.c_int32)]),
    "kat_sampler": (None, [C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, c_d3]),
    "kat_scatter": (C.c_int32, [_P, C.c_int32, C.c_uint64, C.c_uint64, _P, _P, _P, c_d3, c_d3]),
    "kat_path_radiance": (C.c_int32, [_P, C.POINTER(capi.RenderConfig), C.POINTER(C.c_uint64), C.c_int32, c_d3]),
}


def build(force: bool = False) -> str:
    """Compile liboracle_rt.so with the committed Makefile (g++ only, no CUDA)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(["make", "-C", _HERE, "liboracle_rt.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


_api = None


def api() -> capi.Api:
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        _api = capi.Api(lib, "orc_", extra=KAT_SIGNATURES)
    return _api


def new_scene() -> capi.Scene:
    return capi.Scene(api())


def d3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def vec3_op(op, a, b=(0, 0, 0), s=0.0):
    out = (C.c_double * 3)()
    api().kat_vec3(op, d3(a), d3(b), float(s), out)
    return tuple(out)


def path_radiance(scene: capi.Scene, cfg: capi.RenderConfig, path_ids) -> np.ndarray:
    ids = np.ascontiguousarray(path_ids, dtype=np.uint64)
    out = np.zeros((ids.shape[0], 3), dtype=np.float64)
    api().check(api().kat_path_radiance(scene.h, C.byref(cfg), ids.ctypes.data_as(C.POINTER(C.c_uint64)), ids.shape[0],
                                        out.ctypes.data_as(c_d3)))
    return out
