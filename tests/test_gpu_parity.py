"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C-ABI of
librtb200.so; the oracle (oracle/) is only the checker.

E1  intersection parity on deterministic ray batches (prim ids identical, t / p / normal within 1e-5)
E2  image parity at equal spp — sample-exact here, because both sides draw the same Philox streams
E3  unit-level device checks (texture, Perlin, Philox, camera)
plus determinism / sharding / depth / compat-row properties.
"""
import ctypes as C
import math

import numpy as np
import pytest

import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi

import parity_utils as pu

pytestmark = pytest.mark.gpu

SCENES = {
    # id: (box lo, box hi, allowed id-mismatch fraction [, allowed fraction of exact ties])
    13: (-15.0, 15.0, 0.0),     # C1a book-1 classic (spheres, BVH)
    99: (-15.0, 15.0, 0.0),     # C1b as shipped (moving spheres, checker)
    0: (-25.0, 25.0, 0.0),
    3: (-12.0, 12.0, 0.0),      # rect light + spheres + noise texture
    4: (0.0, 555.0, 0.0),       # Cornell box: rects + Translate(RotateY(RectPrism))
    5: (0.0, 555.0, 0.0),       # Cornell smoke, media skipped here (geometric query)
    6: (-600.0, 600.0, 0.0, 0.02),  # book-2 final: boxes, spheres, instance of 1000 spheres, r=5000 shell;
                                    # adjacent ground boxes share faces => exact ties (documented)
    7: (-10.0, 10.0, 0.0),
    9: (-10.0, 10.0, 0.0),      # sphere nested in 20 lists
    10: (-8.0, 8.0, 0.0),       # triangle + sphere
    12: (0.0, 555.0, 0.0),      # triangle inside Cornell walls
    14: (-30.0, 56.0, 2e-4),    # mesh room (f32 triangle storage => documented grazing cases)
}


@pytest.mark.parametrize("scene_id", sorted(SCENES))
def test_E1_intersection_parity(orc, scene_id):
    lo, hi, frac = SCENES[scene_id][:3]
    ties = SCENES[scene_id][3] if len(SCENES[scene_id]) > 3 else 0.0
    g, o = pu.build_pair(orc, scene_id, param=48 if scene_id == 14 else 0)
    cam = pu.camera_fields(orc, o)
    tr = (cam["time1"], cam["time2"])
    batches = {
        "primary": pu.primary_rays(cam, 160, 90),
        "random": pu.random_rays(60000, lo, hi, seed=scene_id + 1, time_range=tr),
    }
    total = 0
    tfrac = 1e-3 if scene_id == 14 else 0.0
    for name, rays in batches.items():
        hg, ho = g.trace_batch(rays), o.trace_batch(rays)
        r = pu.assert_parity(hg, ho, f"scene {scene_id} {name}", max_id_frac=frac, max_tie_frac=ties, max_t_frac=tfrac, rays=rays)
        total += r["hits"]
        sec = pu.secondary_rays(ho, seed=7, time=0.5 * (tr[0] + tr[1]))
        if sec.shape[0]:
            pu.assert_parity(g.trace_batch(sec), o.trace_batch(sec), f"scene {scene_id} {name} secondary", max_id_frac=max(frac, 2e-5),
                             max_tie_frac=ties, max_t_frac=tfrac, rays=sec, require_hits=False)
    assert total > 1000


def test_E1_interval_ends_and_ties(orc):
    # inclusive interval ends (hit.rs:216-219) and "later list element wins" (hit.rs:675-683)
    def build(s):
        a = s.xz_rect(-100, 100, -100, 100, 55, s.metal((1, 1, 1), 0.0))
        b = s.xz_rect(-100, 100, -100, 100, 55, s.diffuse_light((4, 4, 4)))
        sp = s.sphere((0, 0, 0), 4.0, s.lambertian((0.5, 0.5, 0.5)))
        s.set_root(s.list([a, sp, b]))
        s.set_camera((0, 0, 20), (0, 0, 0), (0, 1, 0), 40, 1.0, 0.0, 10, 0, 1)
        s.commit()
    g, o = rtb.new_scene(), orc.new_scene()
    build(g); build(o)
    rays = capi.make_rays([(0.1, 5, 0.2), (0, 0, 20), (0, 0, 20)], [(0.1, 1, 0.2), (0, 0, -1), (0, 0, -4)])
    hg, ho = g.trace_batch(rays), o.trace_batch(rays)
    assert list(hg["prim_id"]) == list(ho["prim_id"]) == [2, 1, 1]
    assert np.array_equal(hg["t"], ho["t"])
    for tmax in (16.0, 15.999999):
        hg, ho = g.trace_batch(rays[1:2], 0.001, tmax), o.trace_batch(rays[1:2], 0.001, tmax)
        assert hg["prim_id"][0] == ho["prim_id"][0]


def test_E1_shared_objects_and_nested_chains(orc):
    # Arc sharing (one object reached through several wrappers), wrapper chains of depth 3, a rotated
    # instance that is NOT wrapped in a Translate (so RotateY's object-space face-forwarding quirk,
    # hit.rs:921, is visible in the normal), a BVH of transformed children and a general medium boundary.
    def build(s):
        m = s.lambertian((0.5, 0.5, 0.5))
        glass = s.dielectric(1.5)
        sph = s.sphere((0, 0, 0), 1.0, glass)
        box = s.box((-1, -1, -1), (1, 2, 1.5), m)
        tri = s.triangle((0, 0, 0), (2, 0, 0.5), (0, 2, 1), m)
        group = s.list([sph, s.translate((3, 0, 0), box)])
        a = s.translate((-6, 0, 0), s.rotate_y(30, s.translate((0, 1, 0), group)))
        b = s.rotate_y(-50, s.list([box, tri]))                        # bare RotateY
        c = s.translate((6, 0, 2), group)                              # the same group again (shared)
        d = s.bvh([s.translate((0, 5, 0), sph), s.translate((0, -5, 0), box), s.sphere((0, 0, 8), 2.0, m)], 0, 1)
        med = s.constant_medium((1, 1, 1), 0.3, s.list([s.sphere((0, 0, -8), 2.0, m), s.sphere((1, 0, -8), 2.0, m)]))  # general boundary
        s.set_root(s.list([a, b, c, d, med]))
        s.set_camera((0, 3, 25), (0, 0, 0), (0, 1, 0), 40, 1.5, 0.0, 10, 0, 1)
        s.commit()
    g, o = rtb.new_scene(), orc.new_scene()
    build(g); build(o)
    assert g.num_prims() == o.num_prims() == 1 + 6 + 1 + 1 + 1 + 2
    cam = pu.camera_fields(orc, o)
    for rays in (pu.primary_rays(cam, 200, 130), pu.random_rays(80000, -12.0, 12.0, seed=21)):
        hg, ho = g.trace_batch(rays), o.trace_batch(rays)
        r = pu.assert_parity(hg, ho, "shared/nested")
        assert r["hits"] > 3000
        assert len(set(np.unique(ho["prim_id"])) - {-1}) >= 6
        hg = g.trace_batch(rays, flags=capi.RT_TRACE_SEEDED_MEDIA, seed=5)
        ho = o.trace_batch(rays, flags=capi.RT_TRACE_SEEDED_MEDIA, seed=5)
        pu.assert_parity(hg, ho, "shared/nested + general medium")
        assert (ho["prim_id"] == 9).sum() > 10  # the medium itself produced records
    # and the render goes through the general-medium kernels (wavefront and fused) identically
    cfg = capi.make_config(60, 1.5, 6, 20, seed=8)
    sg, ag, stg = g.render(cfg, want_accum=True)
    so, ao, sto = o.render(cfg, want_accum=True)
    assert (np.abs(ag - ao) > 1e-5 * np.maximum(np.abs(ao), 2.0 ** 22)).any(axis=2).mean() < 0.02
    _, af, _ = g.render(capi.make_config(60, 1.5, 6, 20, seed=8, flags=8), want_accum=True)
    assert np.array_equal(ag, af)


def test_documented_divergence_rotate_y_bbox_inside_bvh(orc):
    # DESIGN.md divergence 4: RotateY::new computes the rotated corner bounds and then stores the UN-rotated
    # box (hit.rs:886), so inside a reference BvhNode the corners of a rotated object that stick out of
    # its un-rotated box are culled away.  The GPU path bounds the rotated object correctly: wherever the
    # two disagree (beyond edge ties), the GPU sees the box and the reference sees through it.
    def build(s):
        m = s.lambertian((0.5, 0.5, 0.5))
        s.set_root(s.bvh([s.rotate_y(45, s.box((-2, -1, -2), (2, 1, 2), m))], 0, 1))  # single-object node: bbox = the un-rotated box
        s.set_camera((1, 4, 12), (0, 0, 0), (0, 1, 0), 40, 1.5, 0.0, 10, 0, 1)
        s.commit()
    g, o = rtb.new_scene(), orc.new_scene()
    build(g); build(o)
    rays = pu.primary_rays(pu.camera_fields(orc, o), 300, 200)
    hg, ho = g.trace_batch(rays), o.trace_batch(rays)
    diff = hg["prim_id"] != ho["prim_id"]
    assert 50 < diff.sum() < 0.1 * rays.shape[0]
    assert np.all(hg["prim_id"][diff] >= 0) and np.all(hg["prim_id"][diff] < 6)  # the GPU hits a side of the box ...
    assert np.all(ho["prim_id"][diff] == -1)                                      # ... the reference sees through its corners
    px = hg["p"][diff]
    assert np.all(np.maximum(np.abs(px[:, 0]), np.abs(px[:, 2])) > 2.0 - 1e-9)     # only outside the un-rotated box
    same = ~diff & (ho["prim_id"] >= 0)
    assert same.sum() > 1000 and np.allclose(hg["t"][same], ho["t"][same], rtol=1e-12)


def test_empty_world_and_edge_configs(orc):
    # an empty HittableList never hits (hit.rs:660-690): every pixel is the background
    g = rtb.new_scene()
    g.set_root(g.list([]))
    g.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 40, 1.0, 0.0, 10, 0, 1)
    g.set_background((0.25, 0.5, 1.0))
    g.commit()
    scr, acc, st = g.render(capi.make_config(16, 1.0, 3, 5, seed=1), want_accum=True)
    assert st["segments"] == st["paths"] == 16 * 16 * 3
    assert np.all(scr == np.floor(255.9 * np.sqrt([0.25, 0.5, 1.0])))
    h = g.trace_batch(capi.make_rays([(0, 0, 0)], [(0, 0, 1)]))
    assert h["prim_id"][0] == -1
    assert g.trace_batch(np.zeros(0, dtype=capi.RAY_DTYPE)).shape == (0,)
    # Config::new asserts (world.rs:36-40) and bad sample ranges surface as errors, not crashes
    for bad in (capi.make_config(0, 1.0, 1, 1), capi.make_config(8, 1.0, 0, 1), capi.make_config(8, 1.0, 1, 0),
                capi.make_config(8, 1.0, 4, 5, sample_begin=3, sample_end=2), capi.make_config(8, 1.0, 4, 5, sample_end=9)):
        with pytest.raises(capi.RtError):
            g.render(bad)
    # 1 spp, depth 1, odd sizes: exact path count, background only where nothing is hit
    g2, o2 = pu.build_pair(orc, 4)
    sg, ag, stg, so, ao, sto = render_pair(g2, o2, 37, 1.0, 1, 1)
    assert stg["paths"] == sto["paths"] == 37 * 37 and np.array_equal(sg, so)


@pytest.mark.parametrize("scene_id", [5, 6])
def test_seeded_media_parity(orc, scene_id):
    g, o = pu.build_pair(orc, scene_id)
    lo, hi = SCENES[scene_id][:2]
    rays = pu.random_rays(40000, lo, hi, seed=3, time_range=(0.0, 1.0))
    hg = g.trace_batch(rays, flags=capi.RT_TRACE_SEEDED_MEDIA, seed=11)
    ho = o.trace_batch(rays, flags=capi.RT_TRACE_SEEDED_MEDIA, seed=11)
    r = pu.assert_parity(hg, ho, f"media scene {scene_id}", max_tie_frac=0.02)
    medium_ids = {2405, 2408} if scene_id == 6 else {6, 13}
    n_medium = int(np.isin(ho["prim_id"], list(medium_ids)).sum())
    assert n_medium > 200, (n_medium, r)


def unit_op(api, scene, op, ia=0, ib=0, ic=0, idd=0, vals=()):
    inp = (C.c_double * 8)(*list(vals) + [0.0] * (8 - len(vals)))
    out = (C.c_double * 8)()
    api.lib.rt_unit_op.restype = C.c_int32
    api.check(api.lib.rt_unit_op(C.c_void_p(scene.h), op, ia, ib, ic, idd, inp, out))
    return list(out)


def test_E3_unit_ops(orc):
    api = rtb.load()

    def build(s):
        chk = s.tex_checker(s.tex_solid((0.2, 0.3, 0.1)), s.tex_solid((0.9, 0.9, 0.9)))
        noise = s.tex_noise(0.1, seed=7)
        rng = np.random.default_rng(5)
        img = s.tex_image(rng.integers(0, 256, size=(16, 32, 3)).astype(np.float64))
        s.set_root(s.list([s.sphere((0, 0, 0), 1, s.lambertian(chk))]))
        s.set_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20, 16 / 9, 0.1, 10, 0, 10)
        s.commit()
        return chk, noise, img
    g, o = rtb.new_scene(), orc.new_scene()
    chk, noise, img = build(g)
    assert (chk, noise, img) == build(o)
    rng = np.random.default_rng(9)
    out = (C.c_double * 3)()
    nz, tb = C.c_double(), C.c_double()
    for p in rng.uniform(-300, 300, size=(40, 3)):
        u, v = rng.uniform(-0.1, 1.1, size=2)
        for tex in (chk, noise, img):
            orc.api().kat_texture_value(o.h, tex, u, v, (C.c_double * 3)(*p), out)
            got = unit_op(api, g, 0, tex, vals=[u, v, *p])
            np.testing.assert_allclose(got[:3], list(out), atol=2e-6, rtol=0)
        orc.api().kat_perlin(o.h, noise, (C.c_double * 3)(*p), C.byref(nz), C.byref(tb))
        got = unit_op(api, g, 1, 0, vals=list(p))
        assert got[0] == pytest.approx(nz.value, abs=1e-12) and got[1] == pytest.approx(tb.value, abs=1e-12)
    # Perlin lattice zeros (perlin.rs:97-101)
    assert unit_op(api, g, 1, 0, vals=[3.0, -2.0, 7.0])[:2] == [0.0, 0.0]
    # Philox-4x32-10 known answers
    assert unit_op(api, g, 2, 0, 0, 0, 0, vals=[0, 0])[:4] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert unit_op(api, g, 2, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, vals=[0xa4093822, 0x299f31d0])[:4] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # camera rays: same draws, same ray (camera.rs:59-71)
    ray = np.zeros(1, dtype=capi.RAY_DTYPE)
    for pid in (0, 1, 12345, 2 ** 33 + 5):
        orc.api().check(orc.api().kat_camera_ray(o.h, 77, pid, 10, 20, 800, 450, ray.ctypes.data))
        got = unit_op(api, g, 3, 77, 0, pid & 0xffffffff, pid >> 32, vals=[10, 20, 800, 450])
        np.testing.assert_allclose(got[:3], ray["o"][0], atol=1e-14)
        np.testing.assert_allclose(got[3:6], ray["d"][0], atol=1e-13)
        assert got[6] == pytest.approx(ray["time"][0], abs=1e-15)


def render_pair(g, o, W, aspect, spp, depth, seed=5, **kw):
    cfg = capi.make_config(W, aspect, spp, depth, seed=seed, **kw)
    sg, ag, stg = g.render(cfg, want_accum=True)
    so, ao, sto = o.render(cfg, want_accum=True)
    return sg, ag, stg, so, ao, sto


@pytest.mark.parametrize("scene_id,W,aspect,spp", [(13, 96, 1.5, 8), (99, 96, 16 / 9, 8), (4, 64, 1.0, 8), (5, 64, 1.0, 8), (6, 64, 1.0, 6),
                                                    (3, 64, 16 / 9, 8), (1, 64, 16 / 9, 4), (2, 64, 16 / 9, 4), (8, 64, 1.5, 4), (14, 64, 1.0, 4)])
def test_E2_render_matches_oracle_sample_by_sample(orc, scene_id, W, aspect, spp):
    g, o = pu.build_pair(orc, scene_id, param=32 if scene_id == 14 else 0)
    sg, ag, stg, so, ao, sto = render_pair(g, o, W, aspect, spp, 50)
    assert sg.shape == so.shape
    assert stg["paths"] == sto["paths"]
    fg, fo = ag / capi.ACCUM_SCALE, ao / capi.ACCUM_SCALE
    rel = np.abs(fg - fo) / np.maximum(1e-3, np.abs(fo))
    pix_bad = (rel > 1e-5).any(axis=2)
    # identical Philox streams => identical paths except where f32 slabs / FMA contraction flips a
    # discrete decision; those are rare and unbiased
    assert pix_bad.mean() < 0.02, (pix_bad.mean(), stg, sto)
    assert abs(fg.mean() - fo.mean()) <= 0.01 * max(fo.mean(), 1e-3)
    assert abs(stg["segments"] - sto["segments"]) <= 0.005 * sto["segments"]
    q_bad = (np.abs(sg - so) > 1).any(axis=2).mean()
    assert q_bad < 0.02


def _psnr(a, b):
    mse = float(np.mean((a - b) ** 2))
    return 99.0 if mse == 0.0 else 10.0 * math.log10(255.0 ** 2 / mse)


@pytest.mark.parametrize("scene_id,W,aspect,spp", [(13, 120, 1.5, 32), (5, 96, 1.0, 48), (6, 96, 1.0, 32)])
def test_E2_statistical_image_parity(orc, scene_id, W, aspect, spp):
    # SURVEY.md Appendix E2 / BASELINE north_star "images at equal spp within a stated RMSE/PSNR bound": with DIFFERENT seeds on
    # the two sides the quantised 8-bit images (the reference's output space, vec3.rs:89-107) must be as close as two
    # independent renders of the oracle are to each other: PSNR(GPU, oracle A) >= PSNR(oracle B, oracle A) - 0.5 dB,
    # RMSE within 6 % of the noise floor, mean colour within 0.5/255 + 4 sigma and mean linear radiance within 4 sigma.
    g, o = pu.build_pair(orc, scene_id)
    so_a, ao_a, _ = o.render(capi.make_config(W, aspect, spp, 50, seed=101), want_accum=True)
    so_b, ao_b, _ = o.render(capi.make_config(W, aspect, spp, 50, seed=202), want_accum=True)
    sg_c, ag_c, _ = g.render(capi.make_config(W, aspect, spp, 50, seed=303), want_accum=True)
    floor, got = _psnr(so_b, so_a), _psnr(sg_c, so_a)
    assert got >= floor - 0.5, (scene_id, got, floor)
    rmse_floor, rmse_got = math.sqrt(np.mean((so_b - so_a) ** 2)), math.sqrt(np.mean((sg_c - so_a) ** 2))
    assert rmse_got <= 1.06 * rmse_floor, (rmse_got, rmse_floor)
    npix = so_a.shape[0] * so_a.shape[1]
    sig8 = (so_b - so_a).std(axis=(0, 1)) / math.sqrt(2.0 * npix)  # standard error of a channel mean, from the two oracle renders
    assert (np.abs(sg_c.mean(axis=(0, 1)) - so_a.mean(axis=(0, 1))) <= 0.5 + 4.0 * math.sqrt(2.0) * sig8).all()
    # bias check on linear radiance: the GPU mean must sit within 4 sigma of the oracle mean, sigma estimated from the two oracle runs
    la, lb, lc = (x / capi.ACCUM_SCALE / spp for x in (ao_a, ao_b, ag_c))
    sigma = np.abs(la.mean(axis=(0, 1)) - lb.mean(axis=(0, 1))) / math.sqrt(2.0) + la.std(axis=(0, 1)) / math.sqrt(la.shape[0] * la.shape[1])
    assert (np.abs(lc.mean(axis=(0, 1)) - 0.5 * (la + lb).mean(axis=(0, 1))) <= 4.0 * sigma + 1e-4).all()


def test_depth_limit_and_background(orc):
    g, o = pu.build_pair(orc, 13)
    for depth in (1, 2, 5):
        sg, ag, stg, so, ao, sto = render_pair(g, o, 64, 1.5, 4, depth)
        assert stg["segments"] == sto["segments"]
        assert (np.abs(ag - ao) > 1e-5 * np.maximum(ao, 2 ** 20)).mean() < 0.01
    # depth 1: only primary rays that miss everything see the background (world.rs:64-89)
    cfg = capi.make_config(64, 1.5, 2, 1, seed=1)
    _, ag, st = g.render(cfg, want_accum=True)
    assert st["segments"] == st["paths"]
    vals = np.unique(ag.reshape(-1, 3), axis=0) / capi.ACCUM_SCALE
    for v in vals:
        k = round(v[2] / 1.0)  # background blue = 1.0 per sample
        np.testing.assert_allclose(v, np.array([0.7, 0.8, 1.0]) * k, atol=1e-6)


def test_determinism_wave_size_and_sharding(orc):
    api = rtb.load()
    api.lib.rt_scene_set_tuning.restype = C.c_int32
    g = rtb.new_scene()
    g.world_build(5, 1)
    g.commit()
    cfg = capi.make_config(80, 1.0, 12, 50, seed=9)
    _, a1, st1 = g.render(cfg, want_accum=True)
    _, a2, _ = g.render(cfg, want_accum=True)
    assert np.array_equal(a1, a2)  # bit-reproducible (integer accumulation, counter-based RNG)
    api.lib.rt_scene_set_tuning(C.c_void_p(g.h), 4096)
    _, a3, st3 = g.render(cfg, want_accum=True)
    assert np.array_equal(a1, a3) and st3["iterations"] > st1["iterations"]
    api.lib.rt_scene_set_tuning(C.c_void_p(g.h), 1 << 20)
    # sample-range shards (the multi-GPU split) add up exactly
    parts = []
    for lo, hi in ((0, 5), (5, 6), (6, 12)):
        c = capi.make_config(80, 1.0, 12, 50, seed=9, sample_begin=lo, sample_end=hi)
        parts.append(g.render(c, want_accum=True)[1])
    assert np.array_equal(a1, parts[0] + parts[1] + parts[2])
    # a different seed gives a different image
    _, a4, _ = g.render(capi.make_config(80, 1.0, 12, 50, seed=10), want_accum=True)
    assert not np.array_equal(a1, a4)


@pytest.mark.parametrize("scene_id", [13, 99, 8, 5, 6, 14])
def test_fused_mode_is_bit_identical_to_wavefront(scene_id):
    # RT_RENDER_FORCE_FUSED (persistent k_mega) and RT_RENDER_FORCE_WAVEFRONT share device functions, Philox
    # streams and the integer accumulator: same image bit for bit, same segment count.  Scene 99: the fused kernel walks
    # motion-interpolated boxes, the wavefront the union-over-the-shutter boxes; scene 8: the GravitySphere variant
    g = rtb.new_scene()
    g.world_build(scene_id, 3, 48 if scene_id == 14 else 0)  # 14: 4608 triangles => the wavefront uses the warp-scheduled k_extend_p
    g.commit()
    res = []
    aspect = {13: 1.5, 99: 16 / 9, 8: 1.5}.get(scene_id, 1.0)
    for flags in (4, 8, 0):
        _, acc, st = g.render(capi.make_config(72, aspect, 6, 50, seed=4, flags=flags), want_accum=True)
        res.append((acc, st["segments"], st["kernel_launches"]))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
    assert np.array_equal(res[0][0], res[2][0])  # and so is whatever RT_MODE_AUTO picks
    assert res[1][2] == 2 and res[0][2] > 2


@pytest.mark.parametrize("scene_id", [13, 99, 8, 5, 6, 14, 4, 1, 9])
def test_pool_mode_is_bit_identical(scene_id):
    # RT_RENDER_FORCE_POOL (k_pool: every warp runs a wavefront of its own in shared memory - slot lists, resumable walks, per-material
    # batches, media sampled in the classify pass, instances walked one after the other): same image, bit for bit, and the same number of
    # world.hit queries as RT_MODE_AUTO's kernel, also at a depth limit that ends paths early and with tile-sharded renders
    g = rtb.new_scene()
    g.world_build(scene_id, 3, 48 if scene_id == 14 else 0)
    g.commit()
    aspect = {13: 1.5, 99: 16 / 9, 8: 1.5, 1: 1.5, 9: 1.5}.get(scene_id, 1.0)
    for depth, extra in ((50, 0), (3, 0), (50, (1 << 16) | (3 << 24))):  # RT_RENDER_TILE_SHARD(1, 3)
        _, a0, s0 = g.render(capi.make_config(88, aspect, 5, depth, seed=4, flags=extra), want_accum=True)
        _, a1, s1 = g.render(capi.make_config(88, aspect, 5, depth, seed=4, flags=16 | extra), want_accum=True)
        assert np.array_equal(a0, a1) and a0.any() and s0["segments"] == s1["segments"] and s1["kernel_launches"] == 2
    g.close()


@pytest.mark.parametrize("scene_id,param", [(14, 48), (6, 0), (13, 0), (99, 0)])
def test_device_lbvh_builder_parity(orc, scene_id, param, monkeypatch):
    # SURVEY.md 8(f) n1 (csrc/cuda/lbvh.cu replaces BvhNode::new, bvh.rs:14-83): the tree built on the device gives
    # the oracle's hits and, bit for bit, the image of the host SAH tree (closest hit is topology independent)
    monkeypatch.setenv("RTB200_BVH_DEVICE_MIN", "16")
    g, g_sah, o = rtb.new_scene(), rtb.new_scene(), orc.new_scene()
    g.set_bvh_builder(1)
    for s in (g, g_sah, o):
        s.world_build(scene_id, 0xB001, param)
        s.commit()
    assert g.host_check()["device_built_prims"] >= 400 and g_sah.host_check()["device_built_prims"] == 0
    lo, hi, frac = SCENES[scene_id][:3]
    ties = SCENES[scene_id][3] if len(SCENES[scene_id]) > 3 else 0.0
    cam = pu.camera_fields(orc, o)
    tfrac = 1e-3 if scene_id == 14 else 0.0
    for name, rays in (("primary", pu.primary_rays(cam, 160, 90)), ("random", pu.random_rays(40000, lo, hi, seed=11, time_range=(cam["time1"], cam["time2"])))):
        hg, ho = g.trace_batch(rays), o.trace_batch(rays)
        pu.assert_parity(hg, ho, f"lbvh scene {scene_id} {name}", max_id_frac=frac, max_tie_frac=ties, max_t_frac=tfrac, rays=rays)
        assert np.array_equal(hg["prim_id"], g_sah.trace_batch(rays)["prim_id"])
    aspect = 1.5 if scene_id == 13 else (16 / 9 if scene_id == 99 else 1.0)
    cfg = capi.make_config(72, aspect, 6, 50, seed=4)
    _, a_dev, st_dev = g.render(cfg, want_accum=True)
    _, a_sah, st_sah = g_sah.render(cfg, want_accum=True)
    assert np.array_equal(a_dev, a_sah) and st_dev["segments"] == st_sah["segments"]
    with pytest.raises(capi.RtError):
        g.set_bvh_builder(7)


@pytest.mark.parametrize("scene_id", [5, 6])
def test_wavefront_wide_walk_is_bit_identical_to_the_pair_walk(orc, scene_id):
    # Cornell smoke / book-2 final: RT_MODE_AUTO renders them with the wavefront, whose two media kernels walk the main world's 4-wide
    # tree since late round 2 (k_extend<..., WIDE>, auto width); the image equals the sibling-pair walk's bit for bit and the
    # oracle's sums within the E2 bar.  book-2 final has two main-world instances (one under RotateY + Translate): one wide root each.
    acc, segs = {}, {}
    cfg = capi.make_config(64, 1.0, 6, 50, seed=5)
    for width in (2, 4, 0):
        g = rtb.new_scene()
        g.world_build(scene_id, 0xB002, 0)
        g.set_bvh_width(width)
        g.commit()
        assert g.bvh_width() == (2 if width == 2 else 4)
        _, acc[width], st = g.render(cfg, want_accum=True)
        segs[width] = st["segments"]
        assert st["iterations"] > 1  # wavefront mode
        g.close()
    assert np.array_equal(acc[2], acc[4]) and np.array_equal(acc[2], acc[0]) and segs[2] == segs[4] == segs[0]
    o = orc.new_scene()
    o.world_build(scene_id, 0xB002, 0)
    o.commit()
    _, a_o, _ = o.render(cfg, want_accum=True)
    rel = np.abs(acc[0] - a_o) / np.maximum(np.abs(a_o), 2.0 ** 32 * 1e-3)
    assert (rel > 1e-5).any(axis=2).mean() < 0.02
    o.close()


@pytest.mark.parametrize("scene_id,param,env", [(13, 0, None), (8, 0, None), (14, 64, None), (14, 64, {"RTB200_MEGA_WAIT": "0"}), (10, 0, None),
                                                (99, 0, None), (7, 0, None)])
def test_wide_bvh_walk_is_bit_identical_to_the_pair_walk(orc, scene_id, param, env, monkeypatch):
    # rt_scene_set_bvh_width: the 4-wide collapse (csrc/host/bvh_wide.hpp) walked by k_mega / k_mega_r (trace_wide) gives, bit for
    # bit, the image of the sibling-pair walk and the oracle's sums (closest hit is topology independent; ties by depth-first id).
    # 14/64 = 8192 triangles: the resumable kernel; with RTB200_MEGA_WAIT=0 plain k_mega; 8 = GravitySpheres (wide only when forced);
    # 99 / 7 = MovingSpheres: the motion form of the wide nodes (mnodes4), only when forced
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    acc, segs = {}, {}
    aspect = 1.0 if scene_id == 14 else (16 / 9 if scene_id == 99 else 1.5)
    cfg = capi.make_config(96, aspect, 6, 50, seed=9)
    for width in (2, 4, 0):
        g = rtb.new_scene()
        g.world_build(scene_id, 0xB001, param)
        g.set_bvh_width(width)
        g.commit()
        _, acc[width], st = g.render(cfg, want_accum=True)
        segs[width] = st["segments"]
        assert st["iterations"] == 1  # fused mode
        g.close()
    assert np.array_equal(acc[2], acc[4]) and np.array_equal(acc[2], acc[0]) and segs[2] == segs[4] == segs[0]
    if scene_id == 13 and env is None:
        # the counting pass (RT_RENDER_COUNT_EVENTS, bench.py's roofline) walks the tree the fused kernel walks: 4 boxes per visit
        g = rtb.new_scene()
        g.world_build(scene_id, 0xB001, param)
        g.commit()
        assert g.bvh_width() == 4
        _, a_cnt, st_cnt = g.render(capi.make_config(96, aspect, 6, 50, seed=9, flags=2), want_accum=True)
        assert np.array_equal(a_cnt, acc[2]) and st_cnt["box_tests"] % 4 == 0 and 0 < st_cnt["box_tests"] < 40 * st_cnt["segments"]
        g.set_bvh_width(2)
        g.commit()
        assert g.bvh_width() == 2
        _, _, st_pairs = g.render(capi.make_config(96, aspect, 6, 50, seed=9, flags=2), want_accum=True)
        assert st_cnt["prim_tests"][0] <= st_pairs["prim_tests"][0]  # queued leaves behind closest_so_far are dropped
        g.close()
    o = orc.new_scene()
    o.world_build(scene_id, 0xB001, param)
    o.commit()
    _, a_o, _ = o.render(cfg, want_accum=True)
    rel = np.abs(acc[4] - a_o) / np.maximum(np.abs(a_o), 2.0 ** 32 * 1e-3)
    assert (rel > 1e-5).any(axis=2).mean() < 0.02
    o.close()


def test_tile_sharding_is_bit_identical_and_matches_oracle(orc):
    # SURVEY.md 8(e) tile sharding: 4-row bands dealt round-robin; H = 53 (not a multiple of 4), 3 shards, both render modes
    from ray_tracing_series_rust_b200 import sharding
    g, o = pu.build_pair(orc, 13)
    W, aspect, spp = 80, 1.5, 5
    for mode in (4, 8):
        _, full, st = g.render(capi.make_config(W, aspect, spp, 50, seed=6, flags=mode), want_accum=True)
        H = full.shape[0]
        assert H == 53
        total, paths = np.zeros_like(full), 0
        for r in range(3):
            fl = mode | sharding.tile_flags(r, 3)
            scr, part, stp = g.render(capi.make_config(W, aspect, spp, 50, seed=6, flags=fl), want_accum=True)
            mine = sharding.tile_rows(H, r, 3)
            other = [j for j in range(H) if j not in mine]
            assert np.array_equal(part[mine], full[mine]) and not part[other].any() and not scr[other].any()
            assert stp["paths"] == len(mine) * W * spp
            total += part
            paths += stp["paths"]
        assert np.array_equal(total, full) and paths == st["paths"]
    # a tile shard combined with a sample range and the compat_threads row limit, against the oracle
    cfg = capi.make_config(W, aspect, spp, 50, seed=6, compat_threads=10, sample_begin=1, sample_end=4, flags=sharding.tile_flags(2, 3))
    _, ag, sg = g.render(cfg, want_accum=True)
    _, ao, so = o.render(cfg, want_accum=True)
    assert sg["paths"] == so["paths"] and sg["segments"] == so["segments"]
    fg, fo = ag / capi.ACCUM_SCALE, ao / capi.ACCUM_SCALE
    assert (np.abs(fg - fo) <= 1e-5 * np.maximum(1e-3, np.abs(fo))).all()
    assert not ag[50:].any()  # rows >= 10 * (53 / 10) stay unrendered (world.rs:1198-1202)


def test_camera_fields_entry_point(orc):
    # rt_scene_set_camera_fields (the call a Rust Camera::flatten makes) == rt_scene_set_camera with Camera::new's arguments
    g, o = pu.build_pair(orc, 99)
    out = (C.c_double * 24)()
    orc.api().check(orc.api().kat_camera(o.h, out))
    g2 = rtb.new_scene()
    g2.world_build(99, 0xB001, 0)
    g2.set_camera_fields(list(out))
    g2.commit()
    cfg = capi.make_config(64, 16 / 9, 4, 50, seed=2)
    _, a1, _ = g.render(cfg, want_accum=True)
    _, a2, _ = g2.render(cfg, want_accum=True)
    assert np.array_equal(a1, a2)


def test_compat_threads_black_rows_and_ppm(orc, tmp_path):
    # world.rs:1198-1202: rows >= threads * (H / threads) are never rendered (top rows of the PPM)
    g, o = pu.build_pair(orc, 13)
    cfg = capi.make_config(60, 1.6, 2, 10, seed=3, compat_threads=11)  # H = 37 -> 33 rows rendered
    sg, _, stg = g.render(cfg)
    so, _, sto = o.render(cfg)
    assert sg.shape == (37, 60, 3)
    assert np.all(sg[33:] == 0) and np.all(so[33:] == 0) and sg[:33].max() > 0
    assert stg["paths"] == sto["paths"] == 33 * 60 * 2
    pa, pb = tmp_path / "g.ppm", tmp_path / "o.ppm"
    capi.write_ppm(rtb.load(), pa, sg)
    capi.write_ppm(orc.api(), pb, sg)
    assert pa.read_bytes() == pb.read_bytes()
    head = pa.read_text().split("\n")
    assert head[:3] == ["P3", "60 37", "255"] and head[3] == "0 0 0"


def test_camera_shutter_recommit_gravity(orc):
    # C5: per-frame shutter window of the GravitySphere tables (hit.rs:346-379)
    g, o = rtb.new_scene(), orc.new_scene()
    for s in (g, o):
        s.world_build(8, 0xB005)
    for frame in (0, 30, 200):
        cam = ((13, 2, 3), (0, 0, 0), (0, 1, 0), 20, 1.5, 0.1, 10, 0.4 * frame, 0.4 * frame + 0.4)
        for s in (g, o):
            s.set_camera(*cam)
            s.commit()
        camf = pu.camera_fields(orc, o)
        rays = pu.primary_rays(camf, 120, 80, time=0.4 * frame + 0.123)
        pu.assert_parity(g.trace_batch(rays), o.trace_batch(rays), f"gravity frame {frame}")
    with pytest.raises(capi.RtError):
        g.trace_batch(pu.primary_rays(camf, 4, 4, time=5.0))  # outside the committed shutter


def test_render_scene_with_time_frames(orc, tmp_path):
    # world.rs:1249-1330 on the GPU: frames of the bouncing animation, sample-exact against the oracle
    g, o = rtb.new_scene(), orc.new_scene()
    for s in (g, o):
        s.world_build(8, 0xB005)
    for frame in (0, 57):
        cfg = capi.make_config(72, 1.5, 4, 50, seed=5 + frame, threads=4)
        sg, stg = g.render_scene_with_time(0.4 * frame, 0.4 * frame + 0.4, tmp_path / f"g{frame}.ppm", cfg)
        so, sto = o.render_scene_with_time(0.4 * frame, 0.4 * frame + 0.4, tmp_path / f"o{frame}.ppm", cfg)
        assert stg["paths"] == sto["paths"] and abs(stg["segments"] - sto["segments"]) <= 0.005 * sto["segments"]
        assert (np.abs(sg - so) > 1).any(axis=2).mean() < 0.02
        assert (tmp_path / f"g{frame}.ppm").read_text().split("\n")[1] == "72 48"
    # the reference's hard-coded frame: 500x500, 500 spp, THREADS = 11 => rows 495..499 never rendered
    scr, st = g.render_scene_with_time(2.0, 2.4, tmp_path / "full.ppm", None)
    assert scr.shape == (500, 500, 3) and st["paths"] == 495 * 500 * 500
    assert np.all(scr[495:] == 0) and scr[:495].min() > 0  # every rendered pixel sees sky, ground or a sphere
    assert (tmp_path / "full.ppm").read_text().startswith("P3\n500 500\n255\n0 0 0\n")


def test_full_size_book1_properties(orc):
    """BASELINE.json configs[0] at full size (800x533, 500 spp, depth 50): size-independent checks —
    shard linearity, path count, and 48 random pixels recomputed path-by-path by the oracle."""
    g, o = pu.build_pair(orc, 13)
    W, aspect, spp = 800, 1.5, 500
    cfg = capi.make_config(W, aspect, spp, 50, seed=1)
    sg, ag, st = g.render(cfg, want_accum=True)
    H = sg.shape[0]
    assert (H, st["paths"]) == (533, 800 * 533 * 500)
    assert 2.0 < st["segments"] / st["paths"] < 6.0
    halves = [g.render(capi.make_config(W, aspect, spp, 50, seed=1, sample_begin=a, sample_end=b), want_accum=True)[1] for a, b in ((0, 250), (250, 500))]
    assert np.array_equal(ag, halves[0] + halves[1])
    rng = np.random.default_rng(2)
    pix = rng.integers(0, W * H, size=48)
    bad = 0
    for p in pix:
        ids = np.arange(spp, dtype=np.uint64) + np.uint64(p) * np.uint64(spp)
        rad = orc.path_radiance(o, cfg, ids).sum(axis=0)
        got = ag[p // W, p % W] / capi.ACCUM_SCALE
        if np.abs(got - rad).max() > 1e-5 * max(rad.max(), 1.0):
            bad += 1
            assert np.abs(got - rad).max() < 0.05 * max(rad.max(), 1.0)
    assert bad <= 6
    # quantisation (vec3.rs:89-107) of the GPU accumulator equals the Screen the library returned
    ref = np.floor(255.9 * np.clip(np.sqrt(ag / capi.ACCUM_SCALE * (1.0 / spp)), 0, 1))
    assert np.array_equal(ref, sg)


@pytest.mark.parametrize("name,scene_id,W,aspect,spp,camera", [
    ("C2 cornell smoke", 5, 600, 1.0, 1000, None),                 # BASELINE configs[1] at full size
    ("C3 book-2 final", 6, 1000, 1.0, 200, None),                  # configs[2]: full image, 200 of the 10 000 spp (the bench runs all of them)
    ("C5 animation frame 100", 8, 800, 1.5, 200, ((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0, 40.0, 40.4)),  # configs[4], one frame
])
def test_full_size_config_properties(orc, name, scene_id, W, aspect, spp, camera):
    """The other BASELINE configs at their full image size: path count, sample-range and tile-shard linearity
    (bit-exact), and random pixels recomputed path by path by the oracle."""
    from ray_tracing_series_rust_b200 import sharding
    g, o = pu.build_pair(orc, scene_id, seed={5: 0xB002, 6: 0xB002, 8: 0xB005}[scene_id], camera=camera)
    cfg = capi.make_config(W, aspect, spp, 50, seed=3)
    sg, ag, st = g.render(cfg, want_accum=True)
    H = sg.shape[0]
    assert st["paths"] == W * H * spp and 2.0 < st["segments"] / st["paths"] < 12.0
    a, b = spp // 3, spp
    parts = [g.render(capi.make_config(W, aspect, spp, 50, seed=3, sample_begin=x, sample_end=y), want_accum=True)[1] for x, y in ((0, a), (a, b))]
    assert np.array_equal(ag, parts[0] + parts[1])
    tiles = [g.render(capi.make_config(W, aspect, spp, 50, seed=3, flags=sharding.tile_flags(r, 2)), want_accum=True)[1] for r in range(2)]
    assert np.array_equal(ag, tiles[0] + tiles[1])
    rng = np.random.default_rng(4)
    bad = 0
    pix = rng.integers(0, W * H, size=24)
    for p in pix:
        ids = np.arange(spp, dtype=np.uint64) + np.uint64(p) * np.uint64(spp)
        rad = orc.path_radiance(o, cfg, ids).sum(axis=0)
        got = ag[p // W, p % W] / capi.ACCUM_SCALE
        if np.abs(got - rad).max() > 1e-5 * max(rad.max(), 1.0):
            bad += 1  # a path that took another branch at a documented tie / f32 slab grazing case: bounded, not systematic
            assert np.abs(got - rad).max() < 0.05 * max(rad.max(), 1.0), (name, int(p))
    assert bad <= 4, (name, bad)
    ref = np.floor(255.9 * np.clip(np.sqrt(ag / capi.ACCUM_SCALE * (1.0 / spp)), 0, 1))
    assert np.array_equal(ref, sg)


def test_full_size_mesh_room_properties():
    """BASELINE configs[3] (871 200 triangles, 1000x1000) without the oracle (its reference-style build of that mesh is too slow
    for a test): the resumable fused kernel, the plain fused kernel and the wavefront agree bit for bit, the device-built
    BVH gives the same image, and sample ranges add up."""
    g = rtb.new_scene()
    g.world_build(14, 0xB004, 660)
    g.commit()
    assert g.host_check()["tris"] == 871200
    W, spp = 1000, 6
    res = {}
    for label, flags in (("auto", 0), ("fused", 8), ("wavefront", 4)):
        _, acc, st = g.render(capi.make_config(W, 1.0, spp, 50, seed=9, flags=flags), want_accum=True)
        res[label] = acc
        assert st["paths"] == W * W * spp
    assert np.array_equal(res["auto"], res["fused"]) and np.array_equal(res["auto"], res["wavefront"])
    halves = [g.render(capi.make_config(W, 1.0, spp, 50, seed=9, sample_begin=x, sample_end=y), want_accum=True)[1] for x, y in ((0, 2), (2, 6))]
    assert np.array_equal(res["auto"], halves[0] + halves[1])
    gl = rtb.new_scene()
    gl.set_bvh_builder(1)
    gl.world_build(14, 0xB004, 660)
    gl.commit()
    assert gl.host_check()["device_built_prims"] >= 871200
    _, acc_l, _ = gl.render(capi.make_config(W, 1.0, spp, 50, seed=9), want_accum=True)
    assert np.array_equal(acc_l, res["auto"])

