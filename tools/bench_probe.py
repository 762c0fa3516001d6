"""Where does the e2e step spend its wall time?"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi
api = rtb.load()
g = rtb.new_scene(); g.world_build(13, 0xB001, 0); g.commit()
W, H, spp = 800, 533, 500
accum = torch.zeros((H, W, 3), dtype=torch.int64, device="cuda")
screen = torch.zeros((H, W, 3), dtype=torch.float64, device="cuda")
host = torch.empty((H, W, 3), dtype=torch.float64).pin_memory()
stream = torch.cuda.current_stream()
def T(): torch.cuda.synchronize(); return time.time()
for k in range(4):
    t0 = T(); g.commit(); t1 = T()
    cfg = capi.make_config(W, 1.5, spp, 50, seed=100 + k)
    accum.zero_(); st = capi.Stats()
    api.check(api.render_device(g.h, C.byref(cfg), C.c_void_p(accum.data_ptr()), C.c_void_p(stream.cuda_stream), C.byref(st))); t2 = T()
    api.check(api.resolve_device(C.c_void_p(accum.data_ptr()), C.c_void_p(screen.data_ptr()), W, H, spp, H, C.c_void_p(stream.cuda_stream))); t3 = T()
    host.copy_(screen, non_blocking=True); t4 = T()
    print("commit %.1f render %.1f (ms_device %.1f ms_total %.1f) resolve %.1f d2h %.1f" % (1e3*(t1-t0), 1e3*(t2-t1), st.ms_device, st.ms_total, 1e3*(t3-t2), 1e3*(t4-t3)), flush=True)
for k in range(3):
    t1 = T()
    cfg = capi.make_config(W, 1.5, spp, 50, seed=100 + k)
    accum.zero_(); st = capi.Stats()
    api.check(api.render_device(g.h, C.byref(cfg), C.c_void_p(accum.data_ptr()), C.c_void_p(stream.cuda_stream), C.byref(st))); t2 = T()
    print("no-commit render %.1f (ms_device %.1f ms_total %.1f)" % (1e3*(t2-t1), st.ms_device, st.ms_total), flush=True)
