"""A/B of several builds of librtb200.so inside ONE process on one GPU (not part of the product).

    python tools/ab_multi.py [--scenes book1,mesh,shipped,smoke,book2] [--reps 3] [--env VAR=a,b,c] lib_suffix ...

Each suffix names ray_tracing_series_rust_b200/librtb200<suffix>.so ("" = the default build, written as `-`).  Every library is loaded
side by side with ctypes (own CUDA module, own globals), every scene is committed once per library and then rendered `reps` times,
the libraries interleaved, so clocks / thermals are shared.  Prints the best device time per (scene, library) and whether the int64
accumulators of the libraries are bit-identical.  --env sweeps one RTB200_* knob (read in rt_scene_create) per library as well.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from ray_tracing_series_rust_b200 import capi  # noqa: E402

SCENES = {  # name: (scene id, seed, param, width, aspect, spp, identity spp)
    "book1": (13, 0xB001, 0, 800, 1.5, 500, 8),
    "shipped": (99, 0xB001, 0, 800, 1.5, 200, 8),
    "mesh": (14, 0xB004, 660, 1000, 1.0, 20, 2),
    "smoke": (5, 0xB002, 0, 600, 1.0, 200, 8),
    "book2": (6, 0xB002, 0, 1000, 1.0, 50, 4),
    "anim": (8, 0xB005, 240, 800, 1.5, 200, 8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="book1,mesh")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--env", default=None)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    var, vals = (None, [None])
    if a.env:
        var, v = a.env.split("=", 1)
        vals = v.split(",")
    pkg = os.path.join(ROOT, "ray_tracing_series_rust_b200")
    variants = []  # (label, api, env value)
    for sfx in a.libs:
        sfx = "" if sfx == "-" else sfx
        path = os.path.join(pkg, f"librtb200{sfx}.so")
        api = capi.Api(C.CDLL(path), "rt_", extra=capi.DEVICE_SIGNATURES)
        for v in vals:
            variants.append((f"{sfx or 'default'}" + (f" {var}={v}" if var else ""), api, v))
    for name in a.scenes.split(","):
        sid, seed, param, W, aspect, spp, id_spp = SCENES[name]
        scenes = []
        for label, api, v in variants:
            if var:
                os.environ[var] = v
            s = capi.Scene(api)
            t0 = time.time()
            s.world_build(sid, seed, param)
            s.commit()
            scenes.append(s)
            print(json.dumps({"scene": name, "lib": label, "build_commit_s": round(time.time() - t0, 3)}), flush=True)
        ref = None
        same = []
        for s in scenes:  # identity + warm-up
            _, acc, _ = s.render(capi.make_config(W, aspect, id_spp, 50, seed=7), want_accum=True)
            if ref is None:
                ref = acc
            same.append(bool(np.array_equal(ref, acc)))
        best = [1e30] * len(scenes)
        segs = [0] * len(scenes)
        for _ in range(a.reps):
            for i, s in enumerate(scenes):
                _, _, st = s.render(capi.make_config(W, aspect, spp, 50, seed=1))
                best[i] = min(best[i], st["ms_device"])
                segs[i] = st["segments"]
        for i, (label, _, _) in enumerate(variants):
            paths = W * int(W / aspect) * spp
            print(json.dumps({"scene": name, "lib": label, "spp": spp, "best_ms": round(best[i], 3), "Mpaths_s": round(paths / best[i] / 1e3, 1),
                              "vs_first": round(best[0] / best[i], 4), "identical_to_first": same[i], "segments": segs[i]}), flush=True)
        del scenes


if __name__ == "__main__":
    main()
