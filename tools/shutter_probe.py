"""How much of the book-1-as-shipped cost is the union-over-the-shutter bounds of its moving spheres?  Renders the scene with
the shipped shutter [0, 10) and with a collapsed one: the second is the bound a motion-interpolated BVH could reach."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi
for t0, t1 in ((0.0, 10.0), (0.0, 10.0), (0.0, 0.001), (2.5, 2.501), (5.0, 5.001), (9.0, 9.001)):
    g = rtb.new_scene()
    g.world_build(99, 0xB001, 0)
    g.set_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16 / 9, 0.1, 10.0, t0, t1)
    g.commit()
    out = {}
    for flags, key in ((0, "plain"), (3, "counted")):
        _, _, st = g.render(capi.make_config(800, 16 / 9, 100, 50, seed=1, flags=flags))
        out[key] = (round(st["paths"] / st["ms_device"] / 1e3, 1), round(st["box_tests"] / max(st["segments"], 1), 1), round(st["prim_tests"][0] / max(st["segments"], 1), 2))
    print("shutter [%g,%g)" % (t0, t1), "Mpaths/s", out["plain"][0], "boxes/seg", out["counted"][1], "prims/seg", out["counted"][2], flush=True)
