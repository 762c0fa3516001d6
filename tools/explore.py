"""Ad-hoc GPU exploration (not part of the product): times full-size renders of the BASELINE configs."""
import sys, time, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi

def run(scene_id, W, aspect, spp, depth=50, param=0, slots=None, flags=0, label=""):
    g = rtb.new_scene()
    t0 = time.time(); g.world_build(scene_id, 0xB001, param); t1 = time.time(); g.commit(); t2 = time.time()
    if slots:
        import ctypes as C
        rtb.load().lib.rt_scene_set_tuning(C.c_void_p(g.h), slots)
    cfg = capi.make_config(W, aspect, spp, depth, seed=1, flags=flags)
    scr, acc, st = g.render(cfg)
    out = dict(label=label, scene=scene_id, W=W, spp=spp, slots=slots, build_s=round(t1 - t0, 3), commit_s=round(t2 - t1, 3),
               ms_device=round(st["ms_device"], 2), ms_total=round(st["ms_total"], 2), ms_extend=round(st["ms_extend"], 2),
               paths=st["paths"], seg_per_path=round(st["segments"] / max(st["paths"], 1), 3),
               Mpaths_s=round(st["paths"] / st["ms_device"] / 1e3, 2), Mseg_s=round(st["segments"] / st["ms_device"] / 1e3, 2),
               iters=st["iterations"], nodes_per_seg=round(st["box_tests"] / max(st["segments"], 1), 2),
               prims_per_seg=round(st["prim_tests"][0] / max(st["segments"], 1), 2), mean=[round(float(x), 2) for x in scr.mean(axis=(0, 1))])
    print(json.dumps(out), flush=True)
    return scr

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "quick"
    if what == "quick":
        run(13, 800, 1.5, 50, label="C1a warm")
        run(13, 800, 1.5, 500, label="C1a full")
        run(13, 800, 1.5, 50, flags=3, label="C1a counted+timed")
        run(99, 800, 1.5, 100, label="C1b")
        run(5, 600, 1.0, 100, label="C2 smoke")
        run(6, 1000, 1.0, 20, label="C3 final")
        run(14, 1000, 1.0, 10, param=200, label="C4 mesh 80k")
    elif what == "slots":
        for s in (1 << 16, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
            run(13, 800, 1.5, 100, slots=s, label="slots")
    elif what == "modes":
        run(13, 800, 1.5, 50, label="warm")
        for mode in (0, 1):  # 0 = wavefront, 1 = fused
            os.environ["RTB200_MODE"] = str(mode)
            tag = f"mode{mode}"
            run(13, 800, 1.5, 500, label=f"book1 {tag}")
            run(99, 800, 1.5, 200, label=f"book1b {tag}")
            run(5, 600, 1.0, 200, label=f"smoke {tag}")
            run(6, 1000, 1.0, 40, label=f"book2 {tag}")
            run(14, 1000, 1.0, 10, param=660, label=f"mesh871k {tag}")
    elif what == "ab":
        # A/B inside one process (same GPU, same clocks): env var name, values, scenes
        var = sys.argv[2]; vals = sys.argv[3].split(",")
        run(13, 800, 1.5, 50, label="warm")
        for rep in range(2):
            for v in vals:
                os.environ[var] = v
                run(13, 800, 1.5, 500, label=f"book1 {var}={v} #{rep}")
                run(99, 800, 1.5, 200, label=f"book1b {var}={v} #{rep}")
                run(14, 1000, 1.0, 10, param=660, label=f"mesh {var}={v} #{rep}")
                run(5, 600, 1.0, 200, label=f"smoke {var}={v} #{rep}")
                run(6, 1000, 1.0, 50, label=f"book2 {var}={v} #{rep}")
    elif what == "fused":
        run(13, 800, 1.5, 50, label="warm")
        run(13, 800, 1.5, 500, label="book1 final")
        run(99, 800, 1.5, 200, label="book1 shipped")
        run(14, 1000, 1.0, 10, param=660, label="mesh 871k")
    elif what == "bvh":
        # host binned SAH vs device LBVH: commit wall time (with RTB200_COMMIT_TIMING phases on stderr) and the
        # cost of the lower-quality tree when tracing (boxes per segment, paths/s)
        run(13, 800, 1.5, 50, label="warm")
        os.environ["RTB200_COMMIT_TIMING"] = "1"
        for b in ("sah", "lbvh", "sah", "lbvh"):
            os.environ["RTB200_BVH_BUILDER"] = b
            run(14, 1000, 1.0, 20, param=660, label=f"mesh871k {b}")
            run(14, 1000, 1.0, 4, param=660, flags=3, label=f"mesh871k {b} COUNTED")
        os.environ["RTB200_BVH_DEVICE_MIN"] = "16"
        for b in ("sah", "lbvh"):
            os.environ["RTB200_BVH_BUILDER"] = b
            run(13, 800, 1.5, 200, label=f"book1 {b}")
            run(13, 800, 1.5, 50, flags=3, label=f"book1 {b} COUNTED")
            run(6, 1000, 1.0, 40, label=f"book2 {b}")
    elif what == "wave":
        run(13, 800, 1.5, 50, label="warm")
        run(5, 600, 1.0, 300, label="cornell smoke")
        run(6, 1000, 1.0, 100, label="book2 final")
    elif what == "anim":
        run(13, 800, 1.5, 50, label="warm")
        run(8, 800, 1.5, 200, label="bouncing frame (scene 8)")
        run(8, 800, 1.5, 200, label="bouncing frame (scene 8)")
    elif what == "mesh":
        run(14, 1000, 1.0, 4, param=660, label="warm mesh")
        for wait in (16, 20, 24, 0):
            os.environ["RTB200_MEGA_WAIT"] = str(wait)
            run(14, 1000, 1.0, 20, param=660, label=f"mesh wait{wait}")
    elif what == "lbvh_leaf":
        run(14, 1000, 1.0, 4, param=660, label="warm")
        os.environ["RTB200_BVH_BUILDER"] = "lbvh"
        for leaf in (1, 2, 3, 4, 6, 8):
            os.environ["RTB200_MAX_LEAF"] = str(leaf)
            run(14, 1000, 1.0, 20, param=660, label=f"mesh lbvh max_leaf {leaf}")
            run(14, 1000, 1.0, 4, param=660, flags=3, label=f"mesh lbvh max_leaf {leaf} COUNTED")
    elif what == "all":
        run(13, 800, 1.5, 50, label="warm")
        run(13, 800, 1.5, 500, label="book1 final")
        run(99, 800, 1.5, 500, label="book1 shipped")
        run(5, 600, 1.0, 300, label="cornell smoke")
        run(6, 1000, 1.0, 100, label="book2 final")
        run(14, 1000, 1.0, 20, param=660, label="mesh 871k")
        run(5, 600, 1.0, 30, flags=3, label="cornell smoke COUNT")
        run(6, 1000, 1.0, 10, flags=3, label="book2 COUNT")
        run(14, 1000, 1.0, 4, param=660, flags=3, label="mesh COUNT")
