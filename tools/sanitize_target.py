"""Small workload that touches every kernel once, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTB200_BVH_DEVICE_MIN"] = "16"
import numpy as np
import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi, sharding

def render_all_modes(scene_id, seed, param, W, aspect, spp, builder=0):
    g = rtb.new_scene()
    g.set_bvh_builder(builder)
    g.world_build(scene_id, seed, param)
    g.commit()
    out = []
    for flags in (0, 4, 8, 4 | 2, sharding.tile_flags(1, 3)):
        _, acc, st = g.render(capi.make_config(W, aspect, spp, 50, seed=3, flags=flags), want_accum=True)
        out.append((flags, int(acc.sum()), st["segments"]))
    rng = np.random.default_rng(0)
    rays = capi.make_rays(rng.uniform(-10, 10, size=(2048, 3)), rng.normal(size=(2048, 3)))
    h = g.trace_batch(rays)
    print(scene_id, "builder", builder, out, "hits", int((h["prim_id"] >= 0).sum()), flush=True)
    g.close()

render_all_modes(13, 0xB001, 0, 48, 1.5, 2)            # k_mega<..0x1>, k_extend<0,*>, k_shade_all, k_init, k_resolve
render_all_modes(13, 0xB001, 0, 48, 1.5, 2, builder=1)  # lbvh.cu kernels
render_all_modes(99, 0xB001, 0, 48, 16 / 9, 2)          # moving spheres, checker
render_all_modes(8, 0xB005, 0, 48, 1.5, 2)              # gravity spheres
render_all_modes(5, 0xB002, 0, 40, 1.0, 2)              # media fast path, rects + boxes, speculative walk
render_all_modes(6, 0xB002, 0, 40, 1.0, 2)              # book-2: instances, Perlin (shared memory), image texture
render_all_modes(14, 0xB004, 48, 40, 1.0, 2)            # 4608 triangles: k_mega_r, k_extend_p
render_all_modes(14, 0xB004, 48, 40, 1.0, 2, builder=1)
render_all_modes(3, 1, 0, 40, 1.5, 2)                   # general scenes: generic kernels
render_all_modes(4, 1, 0, 40, 1.0, 2)
print("sanitize target done")
