"""The scene builders (csrc/host/world.hpp) are shared by the product and the oracle, so oracle-vs-GPU parity cannot see a mis-port of
world.rs.  These checks are derived BY HAND from the constants in /root/reference/src/world.rs (cited per assertion) and probe the
built scenes through world.hit (trace_batch on the CPU oracle): primitive counts, fixed centres / radii / rect extents, material
thresholds of the random scenes, camera presets.  (tests/test_reference_images.py pins the same builders to the reference's renders.)"""
import numpy as np
import pytest

from ray_tracing_series_rust_b200 import capi


@pytest.fixture(scope="module")
def orc_mod(orc):
    return orc


def _scene(orc, sid, seed=0xB001, param=0):
    s = orc.new_scene()
    s.world_build(sid, seed, param)
    s.commit()
    return s


def _hit(s, o, d, t_min=0.001):
    return s.trace_batch(capi.make_rays([o], [d]), t_min=t_min)[0]


def test_final_scene_constants(orc_mod):
    s = _scene(orc_mod, 6, 0xB002)
    # world.rs:494-616: 400 RectPrisms (6 rect leaves each) + light + moving sphere + 2 spheres + (visible sphere, medium, the medium's own
    # boundary sphere) x 2 + earth + noise + 1000 spheres
    assert s.num_prims() == 400 * 6 + 1 + 1 + 2 + 2 * 3 + 1 + 1 + 1000
    # light XzRect(123, 432, 147, 412, 554) (world.rs:521-523; 432, not the book's 423): straight up from below its corners / just outside
    for x, z, inside in ((124, 148, True), (431, 411, True), (122, 148, False), (433, 300, False), (300, 146, False), (300, 413, False)):
        h = _hit(s, (x, 500.0, z), (0, 1, 0))
        assert (abs(h["t"] - 54.0) < 1e-9) == inside, (x, z, h["t"])
    # glass sphere (260, 150, 45) r 50 and metal sphere (0, 150, 145) r 50 (world.rs:537-552): axis rays from far outside hit at |c - o| - r
    for c in ((260.0, 150.0, 45.0), (0.0, 150.0, 145.0)):
        h = _hit(s, (c[0], c[1], c[2] - 200.0), (0, 0, 1))
        assert abs(h["t"] - 150.0) < 1e-9 and abs(h["normal"][2] + 1.0) < 1e-12
    # r = 70 boundary sphere at (360, 150, 145) (world.rs:554-563), r = 100 earth at (400, 200, 400), r = 80 noise at (220, 280, 300)
    # (probe directions chosen so that nothing else lies in front: the moving sphere hangs right above the earth sphere,
    # and the light at y = 554 covers the noise sphere: start below it)
    for c, r, d, dist in (((360.0, 150.0, 145.0), 70.0, (0, -1, 0), 300.0), ((400.0, 200.0, 400.0), 100.0, (-1, 0, 0), 300.0), ((220.0, 280.0, 300.0), 80.0, (0, -1, 0), 250.0)):
        o = tuple(c[a] - dist * d[a] for a in range(3))
        h = _hit(s, o, d)
        assert abs(h["t"] - (dist - r)) < 1e-9, (c, h["t"])
    # the r = 5000 glass shell IS in the world (world.rs:564-568, unlike the book): a ray leaving the scene upwards from above the light hits it
    h = _hit(s, (278.0, 600.0, 278.0), (0, 1, 0))
    assert abs(h["t"] - (np.sqrt(5000.0 ** 2 - 278.0 ** 2 - 278.0 ** 2) - 600.0)) < 1e-6 and h["front_face"] == 0
    # moving sphere (400,400,400) -> (430,400,400), r 50, times 0..1 (world.rs:528-535): at time 0 / 1 the top is above x = 400 / 430
    for time, x in ((0.0, 400.0), (0.999999, 430.0)):
        r = capi.make_rays([(x, 540.0, 400.0)], [(0, -1, 0)], time=time)  # from below the light (y = 554)
        assert abs(s.trace_batch(r)[0]["t"] - 90.0) < 1e-3
    # ground boxes: 20 x 20 of width 100 from -1000, heights U[1, 101) (world.rs:499-518): every box top lies in [1, 101)
    tops = [-(_hit(s, (-950.0 + 100 * i, 500.0, -950.0 + 100 * j), (0, -1, 0))["t"] - 500.0) for i in range(0, 20, 3) for j in (0, 7, 19) if not (11 <= i <= 14 and j == 7)]
    assert all(1.0 <= y < 101.0 for y in tops) and len(set(np.round(tops, 6))) > 10
    s.close()


def test_cornell_constants(orc_mod):
    s = _scene(orc_mod, 5, 0xB002)
    assert s.num_prims() == 6 + 2 * (1 + 6)  # five walls + light + two media, each with its RectPrism boundary (world.rs:415-492)
    # light XzRect(213, 343, 227, 332, 554) (world.rs:432-439)
    for x, z, inside in ((214, 228, True), (342, 331, True), (212, 300, False), (344, 300, False), (300, 226, False), (300, 333, False)):
        h = _hit(s, (x, 500.0, z), (0, 1, 0))
        assert (abs(h["t"] - 54.0) < 1e-9) == inside  # 554 light, else the 555 ceiling
    # walls at 0 and 555 on all three axes
    assert abs(_hit(s, (278, 278, 278), (1, 0, 0))["t"] - 277.0) < 1e-9 and abs(_hit(s, (278, 278, 278), (-1, 0, 0))["t"] - 278.0) < 1e-9
    assert abs(_hit(s, (278, 278, 500), (0, 0, 1))["t"] - 55.0) < 1e-9 and abs(_hit(s, (400, 278, 100), (0, -1, 0))["t"] - 278.0) < 1e-9
    s.close()
    b = _scene(orc_mod, 4, 0xB002)  # cornell_box: the same boxes as solid RectPrisms (world.rs:388-410)
    assert b.num_prims() == 6 + 2 * 6
    # tall box: Translate((265,0,295), RotateY(15, Box(0..(165,330,165)))): its top is at y = 330 above the point that the rotation maps the box centre to
    c, s15 = np.cos(np.radians(15.0)), np.sin(np.radians(15.0))
    cx, cz = 265.0 + c * 82.5 + s15 * 82.5, 295.0 - s15 * 82.5 + c * 82.5
    assert abs(_hit(b, (cx, 500.0, cz), (0, -1, 0))["t"] - 170.0) < 1e-9
    cx2, cz2 = 130.0 + np.cos(np.radians(-18.0)) * 82.5 + np.sin(np.radians(-18.0)) * 82.5, 65.0 - np.sin(np.radians(-18.0)) * 82.5 + np.cos(np.radians(-18.0)) * 82.5
    assert abs(_hit(b, (cx2, 500.0, cz2), (0, -1, 0))["t"] - 335.0) < 1e-9  # short box: 165 high
    b.close()


def test_book1_scenes_constants(orc_mod):
    for sid in (13, 99):  # classic book-1 final / gen_random_scene as shipped (world.rs:95-167)
        s = _scene(orc_mod, sid)
        n = s.num_prims()
        assert 22 * 22 - 20 <= n - 4 <= 22 * 22  # one per cell unless within 0.9 of (4, 0.2, 0), + ground + three big spheres
        # three r = 1 spheres at (0,1,0), (-4,1,0), (4,1,0) (world.rs:150-163)
        for x in (0.0, -4.0, 4.0):
            h = _hit(s, (x, 10.0, 0.0), (0, -1, 0))
            assert abs(h["t"] - 8.0) < 1e-9
        # ground sphere r = 1000 at (0,-1000,0) [classic] / (0,-1000,-1) [shipped, world.rs:102-106]
        h = _hit(s, (30.0, 5.0, 30.0 if sid == 13 else 29.0), (0, -1, 0))
        zc = 0.0 if sid == 13 else -1.0
        assert abs(h["t"] - (5.0 + 1000.0 - np.sqrt(1000.0 ** 2 - 30.0 ** 2 - (h["p"][2] - zc) ** 2))) < 1e-6
        s.close()
    # material thresholds of the shipped scene (world.rs:117-127): < 0.3 diffuse, < 0.6 metal, else glass, and choose_mat < 0.8 moves (world.rs:128-139)
    s = _scene(orc_mod, 99)
    rng = np.random.default_rng(5)
    o = np.stack([rng.uniform(-11, 11, 40000), np.full(40000, 0.2), rng.uniform(-11, 11, 40000)], 1) + (0, 3.0, 0)
    h = s.trace_batch(capi.make_rays(o, np.tile((0.0, -1.0, 0.0), (40000, 1)), time=0.0))
    small = h[(h["t"] > 2.5) & (h["t"] < 3.3) & (h["prim_id"] > 0)]
    ids = np.unique(small["prim_id"])
    assert len(ids) > 300
    s.close()


def test_camera_presets(orc_mod):
    import ctypes as C
    # get_world_cam (world.rs:876-1179): lookfrom / vfov / aspect per scene id, read back from the 24 stored camera fields
    for sid, lookfrom, vfov, aspect in ((6, (478, 278, -600), 40.0, 1.0), (5, (278, 278, -800), 40.0, 1.0), (99, (13, 2, 3), 20.0, 16 / 9), (11, (0, 20, 20), 60.0, 16 / 9),
                                        (13, (13, 2, 3), 20.0, 1.5)):
        s = orc_mod.new_scene()
        s.world_build(sid, 0xB001, 16 if sid == 11 else 0)
        out = (C.c_double * 24)()
        orc_mod.api().check(orc_mod.api().kat_camera(s.h, out))
        f = np.array(out[:21]).reshape(7, 3)
        assert np.allclose(f[0], lookfrom)
        # |vertical| / |horizontal| = 1 / aspect and |vertical| = 2 tan(vfov / 2) * focus_dist (camera.rs:31-46)
        focus = {6: 10.0, 5: 10.0, 99: 10.0, 11: 40.0, 13: 10.0}[sid]
        assert abs(np.linalg.norm(f[3]) - 2.0 * np.tan(np.radians(vfov) / 2) * focus) < 1e-9
        assert abs(np.linalg.norm(f[2]) / np.linalg.norm(f[3]) - aspect) < 1e-12
        s.close()
