"""Small, fixed workload for ncu: book-1 final scene at reduced spp through rt_render (all kernels)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi

name = sys.argv[1] if len(sys.argv) > 1 else "book1"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfgs = {"book1": (13, 0xB001, 0, 800, 1.5), "shipped": (99, 0xB001, 0, 800, 16 / 9), "smoke": (5, 0xB002, 0, 600, 1.0), "book2": (6, 0xB002, 0, 1000, 1.0), "mesh": (14, 0xB004, 660, 1000, 1.0)}
sid, seed, param, W, aspect = cfgs[name]
g = rtb.new_scene()
g.world_build(sid, seed, param)
g.commit()
scr, _, st = g.render(capi.make_config(W, aspect, spp, 50, seed=1))
print(json.dumps({k: st[k] for k in ("paths", "segments", "iterations", "kernel_launches", "ms_device")}))
