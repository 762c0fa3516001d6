"""Animation driver + progressive accumulation.  CPU part: scheduling, checkpoints, resolve;
GPU part (marked): interrupted + resumed render == one-shot render, frame files, resume skips finished frames."""
import json
import math
import os

import numpy as np
import pytest

import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import animation, capi


def fake_chunk(shape):
    # contribution of sample s to element e: a fixed integer hash (order independent, like the Philox-keyed samples)
    idx = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)

    def f(s0, s1):
        out = np.zeros(shape, dtype=np.int64)
        for s in range(s0, s1):
            out += (idx * 2654435761 + s * 40503) % 1000003
        return out
    return f


def test_progressive_resume_is_exact(tmp_path):
    shape = (5, 7, 3)
    one_shot = fake_chunk(shape)(0, 23)
    ck = str(tmp_path / "still")
    p = animation.ProgressiveRender(fake_chunk(shape), shape, 23, chunk=4, checkpoint=ck, every=2)
    p.run(max_chunks=3)                         # "crash" after 3 chunks: only the checkpoint of chunk 2 is on disk
    assert p.done == 12 and int(np.load(ck + ".npz")["samples_done"]) == 8
    q = animation.ProgressiveRender(fake_chunk(shape), shape, 23, chunk=4, checkpoint=ck, every=2)
    assert q.done == 8 and not q.finished
    acc = q.run()
    assert q.finished and np.array_equal(acc, one_shot)
    z = np.load(ck + ".npz")
    assert int(z["samples_done"]) == 23 and np.array_equal(z["accum"], one_shot)
    assert sorted(os.listdir(tmp_path)) == ["still.npz"]  # ONE file: accumulator and sample count cannot be torn apart
    # a crash while the next checkpoint is being written (temporary present, rename not done) leaves the old pair intact
    q.done, q.accum = 19, fake_chunk(shape)(0, 19)
    real_replace = os.replace
    try:
        os.replace = lambda a, b: (_ for _ in ()).throw(OSError("crash before rename"))
        with pytest.raises(OSError):
            q.save()
    finally:
        os.replace = real_replace
    again = animation.ProgressiveRender(fake_chunk(shape), shape, 23, chunk=4, checkpoint=ck)
    assert again.done == 23 and np.array_equal(again.accum, one_shot)
    os.remove(ck + ".tmp.npz")
    # a finished checkpoint resumes to a no-op; another render's checkpoint is refused
    r = animation.ProgressiveRender(lambda a, b: 1 / 0, shape, 23, chunk=4, checkpoint=ck)
    assert r.finished and np.array_equal(r.run(), one_shot)
    with pytest.raises(ValueError):
        animation.ProgressiveRender(fake_chunk(shape), shape, 24, chunk=4, checkpoint=ck)
    with pytest.raises(ValueError):
        animation.ProgressiveRender(fake_chunk(shape), shape, 0, chunk=4)
    with pytest.raises(ValueError):
        animation.ProgressiveRender(lambda a, b: np.zeros(shape), shape, 4, chunk=4).run()  # wrong dtype


def test_resolve_accumulator_matches_get_normalized_color():
    # vec3.rs:89-107: sqrt(sum / spp), clamp to [0, 1], (255.9 * c) as i32
    spp = 7
    sums = np.array([0.0, 0.5, 7.0, 6.999, 100.0, 1e-9, 3.5])
    acc = np.zeros((1, len(sums), 3), dtype=np.int64)
    acc[0, :, 0] = (sums * 4294967296.0).astype(np.int64)
    out = animation.resolve_accumulator(acc, spp)
    for k, s in enumerate(acc[0, :, 0]):
        c = math.sqrt(float(s) / 4294967296.0 / spp)
        assert out[0, k, 0] == float(int(255.9 * min(max(c, 0.0), 1.0)))
    assert out.max() == 255.0 and out[0, 0, 1] == 0.0
    assert np.array_equal(animation.resolve_accumulator(acc, spp, rendered_rows=0), np.zeros_like(out))


class StubScene:
    def __init__(self):
        self.calls, self.api = [], None

    def set_camera(self, *a):
        self.calls.append(("cam", a[-2], a[-1]))

    def commit(self):
        self.calls.append(("commit",))

    def render(self, cfg):
        self.calls.append(("render", cfg.seed, cfg.samples_per_pixel))
        return np.full((2, 3, 3), float(cfg.seed % 256)), None, {"paths": 6 * cfg.samples_per_pixel}


def test_animation_driver_schedule_and_resume(tmp_path, monkeypatch):
    written = []
    monkeypatch.setattr(capi, "write_ppm", lambda api, path, scr: (written.append(path), open(path, "w").write("P3\n"))[0])
    pat = str(tmp_path / "f%03d.ppm")
    d0 = animation.AnimationDriver(StubScene(), 3, 1.5, 4, 50, pat, rank=0, world=2)
    d1 = animation.AnimationDriver(StubScene(), 3, 1.5, 4, 50, pat, rank=1, world=2)
    assert d0.frames(5) == [0, 2, 4] and d1.frames(5) == [1, 3]
    r0 = d0.run(5)
    assert [f for f, _, st in r0 if st] == [0, 2, 4] and all(os.path.exists(p) for _, p, _ in r0)
    assert d0.scene.calls[:3] == [("cam", 0.0, 0.4), ("commit",), ("render", 5, 4)]          # frame 0: shutter [0, 0.4), seed 5 + f
    assert d0.scene.calls[3:6] == [("cam", 0.8, 1.2000000000000002), ("commit",), ("render", 7, 4)]
    d1.run(5)
    assert sorted(os.listdir(tmp_path)) == ["f%03d.ppm" % f for f in range(5)]
    again = animation.AnimationDriver(StubScene(), 3, 1.5, 4, 50, pat, rank=0, world=1)
    assert all(st is None for _, _, st in again.run(5)) and again.scene.calls == []          # everything already rendered
    os.remove(pat % 3)
    assert [f for f, _, st in again.run(5) if st] == [3]


@pytest.mark.gpu
def test_progressive_gpu_render_equals_one_shot(tmp_path):
    g = rtb.new_scene()
    g.world_build(13, 0xB001, 0)
    g.commit()
    W, aspect, spp, depth = 64, 1.5, 10, 50
    screen, one_shot, _ = g.render(capi.make_config(W, aspect, spp, depth, seed=3), want_accum=True)
    ck = str(tmp_path / "ck")
    p = animation.progressive_scene_render(g, W, aspect, spp, depth, seed=3, chunk=3, checkpoint=ck)
    p.run(max_chunks=2)
    assert p.done == 6 and not np.array_equal(p.accum, one_shot)
    q = animation.progressive_scene_render(g, W, aspect, spp, depth, seed=3, chunk=4, checkpoint=ck)  # resumed with another chunk size
    assert q.done == 6
    assert np.array_equal(q.run(), one_shot)                      # bit-identical to the uninterrupted render
    assert np.array_equal(q.screen(), screen)                     # host resolve == k_resolve


@pytest.mark.gpu
def test_animation_driver_gpu_frames(tmp_path):
    g = rtb.new_scene()
    g.world_build(8, 0xB005, 0)   # gen_random_scene_moving: GravitySpheres
    pat = str(tmp_path / "frame%02d.ppm")
    drv = animation.AnimationDriver(g, 48, 1.5, 4, 50, pat, binary=True)
    res = drv.run(3)
    assert [f for f, _, st in res if st] == [0, 1, 2]
    raws = [open(p, "rb").read() for _, p, _ in res]
    assert all(r.startswith(b"P6\n48 32\n255\n") and len(r) == 13 + 48 * 32 * 3 for r in raws)
    assert raws[0] != raws[2]                                      # the spheres moved
    # the same frame through the reference-shaped per-frame entry gives the same pixels
    scr, _ = g.render_scene_with_time(0.8, 1.2000000000000002, None, capi.make_config(48, 1.5, 4, 50, seed=7))
    px = np.frombuffer(raws[2][13:], dtype=np.uint8).reshape(32, 48, 3)
    assert np.array_equal(px, scr[::-1].astype(np.uint8))
    assert all(st is None for _, _, st in drv.run(3))              # resume: nothing left to do
