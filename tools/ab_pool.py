"""A/B of the pool mode (k_pool, RT_RENDER_FORCE_POOL) against RT_MODE_AUTO's choice, same process, same GPU:
bit-identity of the accumulators on small renders of every scene family, then timings at the BASELINE sizes.
usage: python tools/ab_pool.py [identity|time|sweep] """
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi

POOL = 16


def scene(sid, seed, param, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    s = rtb.new_scene()
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    s.world_build(sid, seed, param)
    s.commit()
    return s


def identity():
    ok = True
    for name, sid, seed, param, W, aspect, spp in (
            ("book1", 13, 0xB001, 0, 240, 1.5, 8), ("book1_shipped", 99, 0xB001, 0, 240, 16 / 9, 6), ("bouncing", 8, 0xB005, 0, 200, 1.5, 6),
            ("mesh8k", 14, 0xB004, 64, 200, 1.0, 8), ("cornell_box", 4, 0xB002, 0, 160, 1.0, 8), ("cornell_smoke", 5, 0xB002, 0, 160, 1.0, 8),
            ("book2", 6, 0xB002, 0, 200, 1.0, 6), ("two_perlin", 1, 0xB001, 0, 160, 1.5, 6), ("simple_light", 3, 0xB001, 0, 160, 1.5, 6),
            ("earth", 2, 0xB001, 0, 120, 1.5, 4), ("nested_lists", 9, 0xB001, 0, 120, 1.5, 4), ("triangle", 10, 0xB001, 0, 120, 1.5, 4),
            ("moving_test", 7, 0xB001, 0, 120, 1.5, 4)):
        s = scene(sid, seed, param)
        for depth in (50, 3):
            a = s.render(capi.make_config(W, aspect, spp, depth, seed=3), want_accum=True)
            b = s.render(capi.make_config(W, aspect, spp, depth, seed=3, flags=POOL), want_accum=True)
            same = bool(np.array_equal(a[1], b[1])) and a[2]["segments"] == b[2]["segments"]
            ok = ok and same
            print(json.dumps({"identity": name, "depth": depth, "bit_identical": same, "segments": [a[2]["segments"], b[2]["segments"]], "nonzero": int((a[1] != 0).sum()),
                              "ndiff": int((a[1] != b[1]).sum())}), flush=True)
        s.close()
    print(json.dumps({"all_bit_identical": ok}), flush=True)
    return 0 if ok else 1


def timed(s, W, aspect, spp, label, flags, reps=2):
    best = None
    for _ in range(reps):
        _, _, st = s.render(capi.make_config(W, aspect, spp, 50, seed=1, flags=flags))
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    print(json.dumps({"label": label, "ms_device": round(best["ms_device"], 3), "Mpaths_s": round(best["paths"] / best["ms_device"] / 1e3, 2),
                      "seg_per_path": round(best["segments"] / best["paths"], 3)}), flush=True)


CONFIGS = (("book1_final", 13, 0xB001, 0, 800, 1.5, 500), ("book1_shipped", 99, 0xB001, 0, 800, 1.5, 200), ("bouncing_frame", 8, 0xB005, 0, 800, 1.5, 200),
           ("mesh871k", 14, 0xB004, 660, 1000, 1.0, 20), ("cornell_smoke", 5, 0xB002, 0, 600, 1.0, 200), ("book2_final", 6, 0xB002, 0, 1000, 1.0, 50))


def times(envs):
    only = os.environ.get("AB_ONLY", "").split(",") if os.environ.get("AB_ONLY") else None
    for name, sid, seed, param, W, aspect, spp in CONFIGS:
        if only and name not in only:
            continue
        s = scene(sid, seed, param)
        s.render(capi.make_config(W, aspect, 4, 50))
        timed(s, W, aspect, spp, f"{name} auto", 0)
        s.close()
        for env in envs:
            s = scene(sid, seed, param, env)
            s.render(capi.make_config(W, aspect, 4, 50, flags=POOL))
            timed(s, W, aspect, spp, f"{name} pool {env}", POOL)
            s.close()
    return 0


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "identity"
    if what == "identity":
        sys.exit(identity())
    if what == "time":
        sys.exit(times([{}]))
    if what == "sweep":
        sys.exit(times([{"RTB200_POOL_SLOTS": "128"}, {"RTB200_POOL_SLOTS": "96"}, {"RTB200_POOL_SLOTS": "64"}, {"RTB200_POOL_SLOTS": "64", "RTB200_POOL_OCC": "5"}]))
