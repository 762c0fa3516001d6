"""Parity pinned to what the REFERENCE itself holds: the three renders it ships (images/book1.png, book2.png,
stanford_dragon.png; README.md:6,18,34), reduced to tests/golden/reference_images.json by
tools/make_reference_image_fixture.py (committed; /root/reference is never read here).

The reference draws from an unseeded thread_rng, so no pixel can be compared one to one.  What the scene CODE fixes
can: the rows that stay black (world.rs:1198-1202 with THREADS = 11), the saturated light, the analytic book-1 sky,
and the mean linear radiance of regions whose content does not depend on the random placement (fog, walls, the big
spheres of final_scene, the three big spheres of book-1).  Bars, as measured when this test was written (oracle at the
test's sizes against the PNGs): sky within 0.2 %, book-1 spheres 1-2 %, dragon-room walls 1-4 %, book-2 regions 2-6 %
(10-12 % for the two spheres that mirror / refract the randomly sized ground boxes).  The asserted bounds below leave
room for Monte-Carlo noise at the test's sample counts.

Both sides are held to the same bars: the CPU oracle here (not gpu), the CUDA path through the C-ABI under -m gpu.
"""
import json
import os

import numpy as np
import pytest

from ray_tracing_series_rust_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_images.json")) as f:
    FIX = json.load(f)

# relative tolerance on the mean linear radiance per region: (oracle at the small CPU size, GPU at full size)
TOL = {
    "book2": {"fog_background": (0.15, 0.08), "moving_sphere": (0.12, 0.08), "blue_medium_sphere": (0.15, 0.10), "noise_sphere": (0.15, 0.12),
              "sphere_cluster": (0.10, 0.06), "ground_boxes": (0.15, 0.12), "glass_sphere": (0.25, 0.20), "metal_sphere": (0.25, 0.20)},
    "stanford_dragon": {"left_wall_green": (0.06, 0.05), "right_wall_blue": (0.06, 0.05), "backdrop_pink": (0.06, 0.05), "mirror_floor": (0.08, 0.06),
                        "below_floor_left": (0.08, 0.07), "below_floor_right": (0.08, 0.07)},
    "book1": {"sky_top": (0.01, 0.01), "sky_upper_right": (0.01, 0.01), "sky_above_horizon": (0.01, 0.01), "ground_far": (0.08, 0.08),
              "ground_near": (0.10, 0.10), "big_lambertian": (0.06, 0.05), "big_metal_top": (0.03, 0.03), "big_glass_low": (0.05, 0.04)},
}


def _region(a, box):
    h, w = a.shape[:2]
    x0, x1, y0, y1 = box
    return a[int(y0 * h):int(y1 * h), int(x0 * w):int(x1 * w)].reshape(-1, a.shape[2])


def _render(scene, width, aspect, spp, seed, compat_threads):
    """-> (8-bit Screen with row 0 = TOP like the PNG, linear radiance per pixel, same orientation)"""
    cfg = capi.make_config(width, aspect, spp, 50, seed=seed, compat_threads=compat_threads, threads=os.cpu_count() or 1)
    scr, acc, st = scene.render(cfg, want_accum=True)
    return scr[::-1], (acc / capi.ACCUM_SCALE / spp)[::-1]


def _check_regions(name, u8, lin, side):
    fx = FIX[name]
    worst = {}
    for rn, r in fx["regions"].items():
        if r["kind"] == "saturated":
            assert r["all_255"] and (_region(u8, r["box"]) == 255).all(), f"{name}/{rn}: the reference's light is saturated everywhere"
            continue
        px = _region(lin, r["box"])
        if r["saturated"]:  # the PNG is clamped at 1 per pixel (mutil.rs:1-9 before the 255.9 scale): do the same where the region touches 255
            px = np.minimum(px, 1.0)
        ours = px.mean(axis=0) if r["kind"] == "mean" else np.median(px, axis=0)
        rel = np.abs(ours / np.array(r["linear_rgb"]) - 1.0).max()
        worst[rn] = float(rel)
        assert rel <= TOL[name][rn][side], f"{name}/{rn}: linear radiance {ours} vs reference image {r['linear_rgb']} (rel {rel:.3f} > {TOL[name][rn][side]})"
    return worst


def _check_black_rows(name, u8, compat_threads):
    """world.rs:1198-1202: the top H - threads * (H / threads) rows are never rendered"""
    H = u8.shape[0]
    n_black = H - compat_threads * (H // compat_threads)
    assert (u8[:n_black] == 0).all(), f"{name}: top {n_black} rows must stay (0,0,0)"
    assert u8[n_black].max() > 0, f"{name}: row {n_black} is rendered"
    return n_black


def _book2(scene, width, spp, side):
    scene.world_build(6, 0xB002, 0)
    scene.commit()
    u8, lin = _render(scene, width, 1.0, spp, 3, 11)
    n_black = _check_black_rows("book2", u8, 11)
    if width == FIX["book2"]["width"]:
        assert list(range(n_black)) == FIX["book2"]["black_rows_from_top"]  # the PNG's ten black rows
    return _check_regions("book2", u8, lin, side)


def _dragon_room(scene, spp, side):
    fx = FIX["stanford_dragon"]
    scene.world_build(11, 0xB004, 48)  # the room of world.rs:681-747 around a small stand-in mesh (the dragon PLY is not shipped)
    scene.commit()
    u8, lin = _render(scene, fx["width"], fx["width"] / fx["height"], spp, 4, 11)
    assert u8.shape[0] == fx["height"]
    n_black = _check_black_rows("stanford_dragon", u8, 11)
    assert list(range(n_black)) == fx["black_rows_from_top"]  # 375 = 11 * 34 + 1: one black row in the PNG
    return _check_regions("stanford_dragon", u8, lin, side)


def _book1(scene, spp, side):
    fx = FIX["book1"]
    scene.world_build(13, 0xB001, 0)
    scene.set_background_gradient((1.0, 1.0, 1.0), (0.5, 0.7, 1.0))  # the sky of the revision that rendered book1.png
    scene.commit()
    u8, lin = _render(scene, fx["width"], 1.5, spp, 5, 0)
    assert u8.shape[0] == fx["height"] and fx["black_rows_from_top"] == []
    # the sky is analytic (primary rays that miss): 8-bit values equal to the PNG's within one level
    ours = np.array([[int(round(float(u8[y, fx["sky_column_x"][0]:fx["sky_column_x"][1], c].mean()))) for c in range(3)] for y in fx["sky_column_rows"]])
    assert np.abs(ours - np.array(fx["sky_column_u8"])).max() <= 1
    return _check_regions("book1", u8, lin, side)


# ---------------------------------------------------------------- the oracle against the reference's images (CPU)
def test_oracle_matches_book2_png():
    import oracle
    _book2(oracle.new_scene(), 250, 24, 0)


def test_oracle_matches_stanford_dragon_png_room():
    import oracle
    _dragon_room(oracle.new_scene(), 8, 0)


def test_oracle_matches_book1_png():
    import oracle
    _book1(oracle.new_scene(), 12, 0)


def test_oracle_constant_background_is_restored():
    """rt_scene_set_background after the gradient switches back to world.rs:86-89's constant"""
    import oracle
    o = oracle.new_scene()
    o.world_build(13, 0xB001, 0)
    o.set_background_gradient()
    o.set_background((0.7, 0.8, 1.0))
    o.commit()
    u8, _ = _render(o, 60, 1.5, 4, 1, 0)
    assert tuple(u8[0, 0]) == (214, 228, 255)  # (255.9 * sqrt((0.7, 0.8, 1.0))) as i32


# ---------------------------------------------------------------- the CUDA path against the reference's images
@pytest.mark.gpu
def test_gpu_matches_book2_png():
    import ray_tracing_series_rust_b200 as rtb
    _book2(rtb.new_scene(), 1000, 200, 1)


@pytest.mark.gpu
def test_gpu_matches_stanford_dragon_png_room():
    import ray_tracing_series_rust_b200 as rtb
    _dragon_room(rtb.new_scene(), 256, 1)


@pytest.mark.gpu
def test_gpu_matches_book1_png():
    import ray_tracing_series_rust_b200 as rtb
    _book1(rtb.new_scene(), 200, 1)


@pytest.mark.gpu
def test_gpu_gradient_sky_equals_oracle_sample_by_sample():
    """the gradient sky on both render modes against the oracle, same seed: same paths, same sums"""
    import oracle
    import ray_tracing_series_rust_b200 as rtb
    g, o = rtb.new_scene(), oracle.new_scene()
    for s in (g, o):
        s.world_build(13, 0xB001, 0)
        s.set_background_gradient()
        s.commit()
    cfg = capi.make_config(96, 1.5, 4, 50, seed=11)
    _, ag, _ = g.render(cfg, want_accum=True)
    _, ao, _ = o.render(cfg, want_accum=True)
    bad = (np.abs(ag - ao) > 1e-5 * np.maximum(capi.ACCUM_SCALE * 1e-3, np.abs(ao))).any(axis=2).mean()
    assert bad < 0.01, bad
    _, aw, _ = g.render(capi.make_config(96, 1.5, 4, 50, seed=11, flags=4), want_accum=True)  # RT_RENDER_FORCE_WAVEFRONT
    assert np.array_equal(ag, aw)


# ---------------------------------------------------------------- the full 871 200-triangle mesh against the oracle
@pytest.mark.gpu
def test_gpu_full_size_mesh_hits_equal_oracle_list_scan():
    """BASELINE configs[3] at full size, ray by ray: the oracle scans the 871 200 Triangle::hit of the mesh as a plain
    HittableList (hit.rs:660-690; closest hit is topology independent, bvh.rs:97-112) - ~4e9 reference triangle tests -
    and the CUDA walk of the SAH tree over f32 vertices must return the same leaves (f32-vertex grazing class bounded as in
    DESIGN.md divergence 7)."""
    import oracle
    import ray_tracing_series_rust_b200 as rtb
    import parity_utils as pu
    g = rtb.new_scene()
    g.world_build(14, 0xB004, 660)
    g.commit()
    assert g.host_check()["tris"] == 871200
    ply = "/tmp/rtb200_mesh_%x_%d.ply" % (0xB004, 660)  # written by world_build(14) (world.hpp stanford_dragon)
    assert os.path.exists(ply)
    o = oracle.new_scene()
    mesh = o.ply_load(ply, 100.0, o.lambertian((0.2, 0.2, 0.2)))  # TriangleModel::load_from_file(..., 100.0).to_hittable(), model.rs:13-76
    light = o.diffuse_light((4, 4, 4))  # material ids in the order of stanford_dragon (world.rs:689-737): mesh, light, then the walls
    room = [o.xy_rect(-100, 100, -100, 100, -20, o.lambertian((0.8, 0.3, 0.3))), o.xy_rect(-100, 100, -100, 100, 20, o.lambertian((1, 1, 1))),
            o.xz_rect(-40, 40, -40, 40, 5, o.metal((0.3, 0.3, 0.3), 0.02)), o.xz_rect(-100, 100, -100, 100, 55, o.metal((1, 1, 1), 0.0)),
            o.yz_rect(-100, 100, -100, 100, -30, o.lambertian((0.3, 0.8, 0.3))), o.yz_rect(-100, 100, -100, 100, 30, o.lambertian((0.3, 0.3, 0.8))),
            o.xz_rect(-100, 100, -100, 100, 55, light)]
    o.set_root(o.list([mesh] + room))  # world.rs:740-747 order; the mesh as a list instead of BvhNode::from_list: same leaves, same ids
    o.set_camera((0, 20, 20), (0, 11, 0), (0, 1, 0), 60.0, 1.0, 0.0, 40.0, 0.0, 10.0)
    o.commit()
    assert o.num_prims() == g.num_prims() == 871200 + 7
    rng = np.random.default_rng(14)
    n = 4096
    # half camera-like rays aimed at the mesh, half rays leaving its surface region in random directions (secondary bounces)
    origin = np.where(np.arange(n)[:, None] < n // 2, np.array([0.0, 20.0, 20.0]), rng.uniform((-11, 5.5, -4.5), (10, 19.5, 4.7), size=(n, 3)))
    target = rng.uniform((-11, 5.5, -4.5), (10, 19.5, 4.7), size=(n, 3))
    d = np.where(np.arange(n)[:, None] < n // 2, target - origin, rng.normal(size=(n, 3)))
    rays = capi.make_rays(origin, d, time=rng.uniform(0, 10, size=n))
    hg, ho = g.trace_batch(rays), o.trace_batch(rays)
    assert (ho["prim_id"] >= 0).mean() > 0.9 and (ho["prim_id"] < 871200).mean() > 0.4  # the batch really exercises the mesh
    # same bars as the small-mesh E1 test: ids identical except < 2e-4 grazing cases, t within 1e-5 relative except < 1e-3 at |cos| < 0.2
    r = pu.assert_parity(hg, ho, "mesh room, 871 200 triangles", max_id_frac=2e-4 + 2.0 / n, max_t_frac=1e-3, rays=rays)
    assert r["hits"] > 0.9 * n
    sec = pu.secondary_rays(ho, seed=8, time=5.0)[:2048]
    pu.assert_parity(g.trace_batch(sec), o.trace_batch(sec), "mesh room, 871 200 triangles, secondary", max_id_frac=2e-4 + 2.0 / sec.shape[0], max_t_frac=1e-3, rays=sec)
