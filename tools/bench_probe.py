"""Why is the event-timed step slower than the wall-timed e2e step?  Times render_device in several settings."""
import sys, os, time, ctypes as C, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi
api = rtb.load()
g = rtb.new_scene(); g.world_build(13, 0xB001, 0); g.commit()
W, H, spp = 800, 533, 500
accum = torch.zeros((H, W, 3), dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step(seed):
    cfg = capi.make_config(W, 1.5, spp, 50, seed=seed)
    accum.zero_(); st = capi.Stats()
    api.check(api.render_device(g.h, C.byref(cfg), C.c_void_p(accum.data_ptr()), C.c_void_p(stream.cuda_stream), C.byref(st)))
    return st
def timed(label, do_flush, n=3):
    for i in range(n):
        if do_flush: flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time(); a.record(stream); st = step(1 + i); b.record(stream); torch.cuda.synchronize(); t1 = time.time()
        print(label, "event_ms %.1f wall_ms %.1f ms_device %.1f ms_total %.1f" % (a.elapsed_time(b), 1e3 * (t1 - t0), st.ms_device, st.ms_total), flush=True)
for i in range(3): step(1)
timed("plain", False)
timed("flush", True)
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
timed("smi100", False)
p.terminate(); p.wait()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-lms", "1000"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
timed("smi1000", False)
p.terminate(); p.wait()
timed("plain2", False)
