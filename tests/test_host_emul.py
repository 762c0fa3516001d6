"""The device functions of csrc/cuda/rt_device.cuh, compiled for the HOST (tests/host_emul, g++ + a small CUDA shim), run
over the host-flattened scene arrays (rt_debug_host_scene) and are compared with the oracle on the E1 ray batches.

This is what a machine without a GPU can say about the device code itself: the flattener, both BVH layouts (sibling
pairs and the 4-wide collapse), the f32 slab tests, the f64 primitive tests, transforms and hit records are the very
source the kernels compile — only the SIMT execution differs (a "warp" is one lane here).  Test infrastructure only:
the product has no CPU path (test_commit_fails_loudly_without_gpu).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi

import parity_utils as pu

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "libemul.so")
    srcs = [os.path.join(HERE, "emul.cpp"), os.path.join(HERE, "cuda_shim.h"),
            os.path.join(ROOT, "ray_tracing_series_rust_b200", "csrc", "cuda", "rt_device.cuh"),
            os.path.join(ROOT, "ray_tracing_series_rust_b200", "csrc", "rt_types.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-I/usr/local/cuda/include", srcs[0], "-o", so])
    lib = C.CDLL(so)
    lib.emul_sizeof_device_scene.restype = C.c_uint64
    lib.emul_trace_batch.restype = C.c_int32
    lib.emul_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int32, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.emul_trace_wide.restype = C.c_int32
    lib.emul_trace_wide.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int32, C.c_void_p]
    if not os.path.exists(rtb.LIB_PATH):
        rtb.build()
    return lib


def host_scene(emul, scene_id, seed=0xB001, param=0, width=2):
    s = rtb.new_scene()
    s.world_build(scene_id, seed, param)
    s.set_bvh_width(width)
    return s, s.debug_host_scene(emul.emul_sizeof_device_scene())


def emul_trace(emul, dscene, rays, t_min=0.001, t_max=float("inf"), flags=capi.RT_TRACE_SKIP_MEDIA, seed=0, wide=0, counts=None):
    rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
    out = np.zeros(rays.shape[0], dtype=capi.HIT_DTYPE)
    cnt = (C.c_uint64 * 2)()
    assert emul.emul_trace_batch(dscene, rays.ctypes.data, rays.shape[0], t_min, t_max, flags, seed, wide, out.ctypes.data, cnt) == 0
    if counts is not None:
        counts.update(boxes=int(cnt[0]), prims=int(cnt[1]))
    return out


def emul_trace_wide(emul, dscene, rays, resume, t_min=0.001, t_max=float("inf")):
    rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
    out = np.zeros(rays.shape[0], dtype=capi.HIT_DTYPE)
    assert emul.emul_trace_wide(dscene, rays.ctypes.data, rays.shape[0], t_min, t_max, resume, out.ctypes.data) == 0
    return out


def oracle_scene(orc, scene_id, seed=0xB001, param=0):
    o = orc.new_scene()
    o.world_build(scene_id, seed, param)
    o.commit()
    return o


# id: (box lo, box hi, allowed id-mismatch fraction, allowed tie fraction) — the bounds of tests/test_gpu_parity.py
SCENES = {
    13: (-15.0, 15.0, 0.0, 0.0),
    99: (-15.0, 15.0, 0.0, 0.0),
    4: (0.0, 555.0, 0.0, 0.0),
    5: (0.0, 555.0, 0.0, 0.0),
    6: (-600.0, 600.0, 0.0, 0.02),
    8: (-15.0, 15.0, 0.0, 0.0),
    14: (-30.0, 56.0, 2e-4, 0.0),
}


@pytest.mark.parametrize("scene_id", sorted(SCENES))
def test_device_functions_on_host_match_the_oracle(orc, emul, scene_id):
    lo, hi, frac, ties = SCENES[scene_id]
    param = 32 if scene_id == 14 else 0
    s, ds = host_scene(emul, scene_id, param=param)
    o = oracle_scene(orc, scene_id, param=param)
    cam = pu.camera_fields(orc, o)
    tr = (cam["time1"], cam["time2"])
    tfrac = 1e-3 if scene_id == 14 else 0.0
    for name, rays in (("primary", pu.primary_rays(cam, 96, 54)), ("random", pu.random_rays(12000, lo, hi, seed=scene_id + 1, time_range=tr))):
        he, ho = emul_trace(emul, ds, rays), o.trace_batch(rays)
        pu.assert_parity(he, ho, f"emul scene {scene_id} {name}", max_id_frac=frac, max_tie_frac=ties, max_t_frac=tfrac, rays=rays)
        sec = pu.secondary_rays(ho, seed=7, time=0.5 * (tr[0] + tr[1]))
        if sec.shape[0]:
            pu.assert_parity(emul_trace(emul, ds, sec), o.trace_batch(sec), f"emul scene {scene_id} {name} secondary", max_id_frac=max(frac, 2e-5),
                             max_tie_frac=ties, max_t_frac=tfrac, rays=sec, require_hits=False)
    s.close()
    o.close()


@pytest.mark.parametrize("scene_id,param", [(14, 32), (14, 96), (13, 0), (10, 0)])
def test_wide_walk_equals_pair_walk(orc, emul, scene_id, param):
    """The 4-wide collapse (bvh_wide.hpp) + trace_wide return the SAME hit records, bit for bit, as the sibling-pair walk:
    closest hit is topology independent and exact ties are decided by depth-first id, not by visiting order."""
    s2, d2 = host_scene(emul, scene_id, param=param, width=2)
    s4, d4 = host_scene(emul, scene_id, param=param, width=4)
    lo, hi = (-30.0, 56.0) if scene_id == 14 else (-15.0, 15.0)
    o = oracle_scene(orc, scene_id, param=param)
    cam = pu.camera_fields(orc, o)
    batches = [pu.primary_rays(cam, 128, 72), pu.random_rays(20000, lo, hi, seed=11, time_range=(cam["time1"], cam["time2"]))]
    batches.append(pu.secondary_rays(o.trace_batch(batches[0]), seed=3, time=cam["time1"]))
    n_hits = 0
    for rays in batches:
        ref = emul_trace(emul, d2, rays)
        for resume in (0, 1):
            w = emul_trace_wide(emul, d4, rays, resume)
            assert w.tobytes() == ref.tobytes(), (scene_id, resume, int((w["prim_id"] != ref["prim_id"]).sum()))
        n_hits += int((ref["prim_id"] >= 0).sum())
    assert n_hits > 1000
    for x in (s2, s4, o):
        x.close()


@pytest.mark.parametrize("scene_id", [4, 5, 6, 8, 13, 14, 99])
def test_wide_world_hit_equals_pair_world_hit(orc, emul, scene_id):
    """world_hit<..., WIDE> (what k_extend<..., WIDE> runs when rt_scene_set_bvh_width(4) forces it): every main-world instance
    through its own 4-wide root, wrappers and media (pair-walked boundaries, seeded draws) included: identical hit records."""
    param = 32 if scene_id == 14 else 0
    s2, d2 = host_scene(emul, scene_id, param=param, width=2)
    s4, d4 = host_scene(emul, scene_id, param=param, width=4)
    o = oracle_scene(orc, scene_id, param=param)
    cam = pu.camera_fields(orc, o)
    lo, hi = SCENES[scene_id][:2]
    prim = pu.primary_rays(cam, 120, 68)
    batches = [prim, pu.random_rays(15000, lo, hi, seed=5, time_range=(cam["time1"], cam["time2"])), pu.secondary_rays(o.trace_batch(prim), seed=3, time=cam["time1"])]
    for rays in batches:
        for flags in (capi.RT_TRACE_SKIP_MEDIA, capi.RT_TRACE_SEEDED_MEDIA):
            c2, c4 = {}, {}
            ref = emul_trace(emul, d2, rays, flags=flags, seed=17, counts=c2)
            w = emul_trace(emul, d4, rays, flags=flags, seed=17, wide=1, counts=c4)
            assert w.tobytes() == ref.tobytes(), (scene_id, flags, int((w["prim_id"] != ref["prim_id"]).sum()))
            assert c4["prims"] <= c2["prims"]  # queued leaves behind closest_so_far are dropped
    for x in (s2, s4, o):
        x.close()


def test_wide_collapse_is_built_only_where_asked_or_measured(emul):
    nbytes = emul.emul_sizeof_device_scene()
    for scene_id, width, built in ((5, 0, False), (6, 0, False), (8, 0, False), (13, 0, True), (13, 2, False), (6, 4, True), (5, 4, True)):
        s, ds = host_scene(emul, scene_id, width=width)
        assert (emul.emul_trace_batch(ds, None, 0, 0.001, 1.0, 0, 0, 1, None, None) == 0) == built, (scene_id, width)
        s.close()
    s = rtb.new_scene()
    with pytest.raises(Exception):
        s.set_bvh_width(3)
    with pytest.raises(Exception):
        s.debug_host_scene(nbytes - 8)
    s.close()
