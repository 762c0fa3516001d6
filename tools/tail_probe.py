"""Fixed cost of one fused render (launch + ramp + tail): ms_device of book-1 final at several spp, linear fit a + b * spp, for two depth limits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi
g = rtb.new_scene(); g.world_build(13, 0xB001, 0); g.commit()
g.render(capi.make_config(800, 1.5, 20, 50))
for depth in (50, 8):
    xs, ys = [], []
    for spp in (16, 31, 62, 63, 125, 250, 500):
        best = min(g.render(capi.make_config(800, 1.5, spp, depth, seed=s))[2]["ms_device"] for s in (1, 2, 3))
        xs.append(spp); ys.append(best)
    b, a = np.polyfit(xs, ys, 1)
    print(f"depth {depth}: " + ", ".join(f"{x} spp {y:.3f} ms" for x, y in zip(xs, ys)) + f" | fit {a:.3f} ms + {b:.5f} ms/spp")
