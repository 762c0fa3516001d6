"""Generates the committed golden fixtures from the CPU oracle (oracle/), in this container.

    python tests/golden/make_golden.py

The reference itself is Rust and cannot run here, so the fixtures are the ORACLE's outputs on fixed seeded
inputs (the oracle is pinned to the reference by tests/test_oracle_kat.py).  They serve two purposes:
the CPU suite checks that the oracle still reproduces them bit for bit (a regression pin of the checker),
and the GPU suite compares the CUDA path against them without trusting a freshly built oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
import parity_utils as pu  # noqa: E402
from ray_tracing_series_rust_b200 import capi  # noqa: E402

SCENES = {13: (-15.0, 15.0), 99: (-15.0, 15.0), 5: (0.0, 555.0), 6: (-600.0, 600.0), 14: (-30.0, 56.0), 8: (-15.0, 15.0), 4: (0.0, 555.0)}


def rays_for(scene_id, o):
    lo, hi = SCENES[scene_id]
    cam = pu.camera_fields(oracle, o)
    a = pu.primary_rays(cam, 24, 16)
    b = pu.random_rays(400, lo, hi, seed=1000 + scene_id, time_range=(cam["time1"], cam["time2"]))
    return np.concatenate([a, b])


def main():
    for sid in SCENES:
        o = oracle.new_scene()
        o.world_build(sid, 0xB001, 24 if sid == 14 else 0)
        o.commit()
        rays = rays_for(sid, o)
        hits = o.trace_batch(rays)
        media = o.trace_batch(rays, flags=capi.RT_TRACE_SEEDED_MEDIA, seed=7)
        cfg = capi.make_config(24, 1.0, 3, 50, seed=11, threads=2)
        scr, acc, st = o.render(cfg, want_accum=True)
        np.savez_compressed(os.path.join(HERE, f"scene{sid}.npz"), rays=rays, hits=hits, hits_media=media, screen=scr.astype(np.uint8), accum=acc,
                            segments=np.int64(st["segments"]))
        print(sid, rays.shape[0], "rays,", int((hits["prim_id"] >= 0).sum()), "hits, segments", st["segments"])


if __name__ == "__main__":
    main()
