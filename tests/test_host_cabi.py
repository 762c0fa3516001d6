"""CPU-side checks of the product library (no GPU needed): the C-ABI loads and exports every symbol
include/rtb200.h declares, the builder validates its arguments like the reference's asserts/panics,
and the host half of rt_scene_commit (flatten + BVH build) satisfies its invariants on every scene
of the reference's scene library.  No compute call is made here; without a device commit/render
must fail loudly with RT_ERR_CUDA (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ray_tracing_series_rust_b200 as rtb
from ray_tracing_series_rust_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def api():
    if not os.path.exists(rtb.LIB_PATH):
        rtb.build()
    return rtb.load()


def test_exports_every_declared_symbol(api):
    hdr = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    names = set(re.findall(r"RTB_FN\((\w+)\)\s*\(", hdr))
    names = {"rt_" + n for n in names} | set(re.findall(r"\b(rt_\w+)\s*\(", hdr))
    names -= {"rt_scene", "rt_status", "rt_render_config", "rt_stats", "rt_ray", "rt_hit", "rt_prim_type", "rt_mat_type"}
    assert len(names) >= 50
    lib = C.CDLL(rtb.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in api.version()


def test_struct_layouts_match_header():
    # rt_render_config / rt_stats / rt_ray / rt_hit as ctypes/numpy see them
    assert C.sizeof(capi.RenderConfig) == 56
    assert C.sizeof(capi.Stats) == 8 * (3 + 8 + 5 + 1 + 2) + 3 * 8
    assert capi.RAY_DTYPE.itemsize == 56 and capi.HIT_DTYPE.itemsize == 88


def test_builder_ids_and_errors(api):
    s = rtb.new_scene()
    t0 = s.tex_solid((0.1, 0.2, 0.3))
    t1 = s.tex_checker(t0, t0)
    assert (t0, t1) == (0, 1)
    m0 = s.lambertian(t1)
    m1 = s.metal((1, 1, 1), 7.0)
    assert (m0, m1) == (0, 1)
    a = s.sphere((0, 0, 0), 1.0, m0)
    b = s.box((0, 0, 0), (1, 1, 1), m1)
    assert (a, b) == (0, 1)
    with pytest.raises(capi.RtError) as e:
        s.sphere((0, 0, 0), 1.0, 99)
    assert e.value.code == -1 and "material" in str(e.value)
    with pytest.raises(capi.RtError):
        s.tex_checker(0, 42)
    with pytest.raises(capi.RtError):
        s.translate((0, 0, 0), 1234)
    with pytest.raises(capi.RtError) as e:
        s.bvh([], 0, 1)  # bvh.rs:27-28: BvhNode over nothing panics in the reference
    assert e.value.code == -6
    with pytest.raises(capi.RtError) as e:
        s.commit()  # no root
    assert e.value.code == -2
    # a medium takes the next texture AND material ids (ConstantMedium::from_color, hit.rs:945-951)
    med = s.constant_medium((1, 1, 1), 0.01, b)
    assert s.tex_solid((0, 0, 0)) == 3 and s.dielectric(1.5) == 3 and med == 2


def test_moving_sphere_needs_a_time_interval(api):
    s = rtb.new_scene()
    m = s.lambertian((0.5, 0.5, 0.5))
    assert s.moving_sphere((0, 0, 0), (0, 1, 0), 0.0, 1.0, 0.5, m) == 0
    with pytest.raises(capi.RtError) as e:
        s.moving_sphere((0, 0, 0), (0, 1, 0), 2.0, 2.0, 0.5, m)   # get_center would divide by zero (hit.rs:275-278)
    assert e.value.code == -1


def test_config_asserts_and_image_height(api):
    cfg = capi.make_config(800, 1.5, 500, 50)
    assert api.image_height(C.byref(cfg)) == 533  # (800 / 1.5) as i32
    assert api.image_height(C.byref(capi.make_config(600, 1.6, 1, 1))) == 375
    assert api.image_height(C.byref(capi.make_config(0, 1.6, 1, 1))) < 0


def host_check(api, scene):
    out = (C.c_int64 * 16)()
    api.check(api.lib.rt_scene_host_check(C.c_void_p(scene.h), out))
    return list(out)


@pytest.mark.parametrize("scene_id,param", [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (7, 0), (8, 0), (9, 0),
                                            (10, 0), (12, 0), (13, 0), (99, 0), (14, 24)])
def test_flatten_and_bvh_invariants(api, scene_id, param):
    api.lib.rt_scene_host_check.restype = C.c_int32
    s = rtb.new_scene()
    s.world_build(scene_id, 0xB001, param)
    st = host_check(api, s)
    nodes, depth, n_main, n_inst, n_media = st[0:5]
    per_type = st[5:11]
    assert st[12] == 0, f"BVH invariant violations: {st}"
    assert depth <= 60 and nodes >= 2 and n_main >= 1
    expect = {
        4: dict(inst=3, media=0, rect=6, box=2),
        5: dict(inst=3, media=2, rect=6, box=2),         # walls + two boundary worlds
        6: dict(inst=4, media=2, box=400, rect=1, moving=1),
        13: dict(inst=1, media=0),
        14: dict(inst=1, media=0, tri=2 * 24 * 24, rect=7),
    }.get(scene_id)
    if expect:
        assert n_inst == expect["inst"] and n_media == expect["media"]
        if "rect" in expect:
            assert per_type[3] == expect["rect"]
        if "box" in expect:
            assert per_type[4] == expect["box"]
        if "tri" in expect:
            assert per_type[5] == expect["tri"]
        if "moving" in expect:
            assert per_type[1] == expect["moving"]
    if scene_id == 6:
        assert per_type[0] == 1000 + 8  # cluster + 6 visible spheres + 2 medium boundaries
        assert st[13] == 400 * 6 + 1 + 1 + 6 + 2 + 2 + 1000  # numbered leaves incl. the two media
    if scene_id == 8:
        assert per_type[2] > 400 and per_type[0] == 4  # gravity spheres + ground + three big spheres


def test_nested_transforms_flatten_to_chains(api):
    api.lib.rt_scene_host_check.restype = C.c_int32
    s = rtb.new_scene()
    m = s.lambertian((0.5, 0.5, 0.5))
    inner = s.list([s.sphere((0, 0, 0), 1, m), s.rotate_y(30, s.box((0, 0, 0), (1, 2, 3), m))])
    root = s.list([s.translate((5, 0, 0), inner), s.sphere((9, 9, 9), 1, m), s.translate((5, 0, 0), inner)])
    s.set_root(root)
    st = host_check(api, s)
    # chains: [], [T1], [T1,R], [T2], [T2,R]  (the two translates are distinct objects)
    assert st[3] == 5 and st[12] == 0
    assert st[13] == 1 + 6 + 1  # shared leaves are numbered once


def test_medium_inside_medium_boundary_is_unsupported(api):
    s = rtb.new_scene()
    m = s.lambertian((0.5, 0.5, 0.5))
    inner = s.constant_medium((1, 1, 1), 0.1, s.sphere((0, 0, 0), 1, m))
    s.set_root(s.constant_medium((1, 1, 1), 0.1, inner))
    out = (C.c_int64 * 16)()
    api.lib.rt_scene_host_check.restype = C.c_int32
    assert api.lib.rt_scene_host_check(C.c_void_p(s.h), out) == -3


def test_gravity_shutter_past_the_table_is_refused(api):
    # GravitySphere::get_center (hit.rs:370-395) leaves its 100 001-entry table for time >= ~100 and re-integrates with other constants;
    # that fallback is not reproduced (DESIGN.md divergence 5), so such a shutter is an error at commit, not a silent clamp
    s = rtb.new_scene()
    s.world_build(8, 0xB005)
    s.set_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0, 99.5, 100.5)
    with pytest.raises(capi.RtError) as e:
        s.commit()
    assert e.value.code == -3 and "hit.rs:381-393" in str(e.value)
    s.set_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0, 95.6, 96.0)  # the last frame of the 240-frame schedule
    if has_cuda():
        s.commit()
    else:
        with pytest.raises(capi.RtError) as e:
            s.commit()
        assert e.value.code == -5  # the scene itself is fine: only the device is missing
    s.close()


def test_ply_and_ppm_io_roundtrip(api, orc, tmp_path):
    # ASCII PLY subset of model.rs:13-62: product loader vs oracle loader on the same file
    ply = tmp_path / "m.ply"
    ply.write_text("ply\nformat ascii 1.0\ncomment x\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                   "element face 2\nproperty list uchar int vertex_indices\nend_header\n"
                   "0 0 0\n1 0 0\n1 1 0.5\n0 1 0 9 9\n3 0 1 2\n3 0 2 3\n")
    s = rtb.new_scene()
    m = s.lambertian((0.2, 0.2, 0.2))
    mesh = s.ply_load(ply, 100.0, m)
    s.set_root(s.bvh([mesh], 0, 1))
    api.lib.rt_scene_host_check.restype = C.c_int32
    st = host_check(api, s)
    assert st[10] == 2 and st[12] == 0 and st[13] == 2
    with pytest.raises(capi.RtError) as e:
        s.ply_load(tmp_path / "missing.ply", 1.0, m)
    assert e.value.code == -4
    bad = tmp_path / "bad.ply"
    bad.write_text("ply\nelement vertex 1\nelement face 1\nend_header\n0 0 0\n3 0 1 2\n")
    with pytest.raises(capi.RtError):
        s.ply_load(bad, 1.0, m)
    # P3 writer: byte-identical between product and oracle, rows reversed (screen.rs:40-48)
    rng = np.random.default_rng(0)
    scr = rng.integers(0, 256, size=(5, 7, 3)).astype(np.float64)
    pa, pb = tmp_path / "a.ppm", tmp_path / "b.ppm"
    capi.write_ppm(api, pa, scr)
    capi.write_ppm(orc.api(), pb, scr)
    ta = pa.read_bytes()
    assert ta == pb.read_bytes()
    lines = ta.decode().split("\n")
    assert lines[0] == "P3" and lines[1] == "7 5" and lines[2] == "255"
    assert lines[3] == " ".join(str(int(x)) for x in scr[4, 0])  # first emitted row = top = Screen row H-1
    # P3 reader (screen.rs:61-95) through Image::from_ppm: both libraries accept the file they wrote
    s2, o2 = rtb.new_scene(), orc.new_scene()
    assert s2.tex_image_ppm(pa) == 0 and o2.tex_image_ppm(pa) == 0
    out = (C.c_double * 3)()
    orc.api().kat_texture_value(o2.h, 0, 0.0, 1.0, (C.c_double * 3)(0, 0, 0), out)  # u=0, v=1 -> file row 0, col 0
    assert tuple(out) == tuple(scr[4, 0] / 255.0)


def test_ascii_ply_number_formats_and_errors(api, tmp_path):
    # the ASCII reader (std::from_chars, body lines parsed on all cores) accepts what Rust's str::parse accepts after split_whitespace
    # (model.rs:44-57: signs, exponents, extra columns, CRLF), gives the binary reader's values bit for bit, and names the first bad line
    hdr = "ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n"
    body = "+1.5 -2.5e-1 3E2\r\n\t0.1  1e-320   -0.0 7 7\r\n 12345678.901234567 .5 5.\n1 1 1\n3 0 1 2\n 3  +1 2 3 \n"
    ply = tmp_path / "fmt.ply"
    ply.write_text(hdr % (4, 2) + body)
    binp = tmp_path / "fmt.bin.ply"
    capi.ply_convert_binary(api, ply, binp)
    raw = binp.read_bytes()
    assert b"property float x" in raw  # the binary twin stores f32 vertices (as the device does)
    vals = np.frombuffer(raw[raw.index(b"end_header\n") + 11:], dtype="<f4", count=12)
    want = np.array([1.5, -0.25, 300.0, 0.1, 1e-320, -0.0, 12345678.901234567, 0.5, 5.0, 1.0, 1.0, 1.0]).astype("<f4")
    assert vals.tobytes() == want.tobytes()
    s = rtb.new_scene()
    m = s.lambertian((0.2, 0.2, 0.2))
    s.ply_load(ply, 2.0, m)
    crlf = tmp_path / "crlf.ply"  # CRLF header lines too (exported from a Windows tool): same mesh
    crlf.write_bytes((hdr % (4, 2) + body).replace("\r\n", "\n").replace("\n", "\r\n").encode())
    s.ply_load(crlf, 2.0, m)
    cases = {
        "short vertex line": (hdr % (2, 0) + "0 0 0\n1 1\n", "bad vertex line"),
        "not a number": (hdr % (1, 0) + "0 zero 0\n", "bad vertex line"),
        # a value that only strtod spells ("inf") on a short line must not be completed from the next line (the reference panics, model.rs:44-48)
        "short line ending in a strtod token": (hdr % (2, 0) + "0 inf\n1 1 1\n", "bad vertex line"),
        "short line, number borrowed across the newline": (hdr % (2, 0) + "0 0\n0x1p0 1 1\n", "bad vertex line"),
        "short face line": (hdr % (3, 1) + "0 0 0\n1 0 0\n0 1 0\n3 0 1\n", "bad face line"),
        "index out of range": (hdr % (3, 1) + "0 0 0\n1 0 0\n0 1 0\n3 0 1 3\n", "out of range"),
        "truncated vertices": (hdr % (3, 1) + "0 0 0\n1 0 0\n", "truncated vertex list"),
        "truncated faces": (hdr % (3, 2) + "0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n", "truncated face list"),
    }
    for name, (text, msg) in cases.items():
        f = tmp_path / "bad.ply"
        f.write_text(text)
        with pytest.raises(capi.RtError) as e:
            s.ply_load(f, 1.0, m)
        assert msg in str(e.value), (name, str(e.value))
    # a body large enough for the parallel path: same arrays as the serial reader's would be (checked through the binary twin)
    rng = np.random.default_rng(3)
    nv, nf = 50000, 40000
    v = rng.normal(size=(nv, 3)) * 10.0 ** rng.integers(-8, 8, size=(nv, 1))
    fidx = rng.integers(0, nv, size=(nf, 3))
    big = tmp_path / "big.ply"
    with open(big, "w") as fh:
        fh.write(hdr % (nv, nf))
        fh.write("".join("%r %r %r\n" % (a, b, c) for a, b, c in v.tolist()))
        fh.write("".join("3 %d %d %d\n" % (a, b, c) for a, b, c in fidx.tolist()))
    bigb = tmp_path / "big.bin.ply"
    capi.ply_convert_binary(api, big, bigb)
    raw = bigb.read_bytes()
    off = raw.index(b"end_header\n") + 11
    assert np.frombuffer(raw[off:], dtype="<f4", count=3 * nv).tobytes() == v.astype("<f4").tobytes()  # repr round-trips: correctly rounded parse
    idx = np.frombuffer(raw[off + 12 * nv:], dtype=np.uint8).reshape(nf, 13)
    assert (idx[:, 0] == 3).all() and np.array_equal(np.ascontiguousarray(idx[:, 1:]).view("<u4").reshape(nf, 3), fidx.astype("<u4"))


def test_binary_io_fast_paths(api, tmp_path):
    # P6 out/in and binary_little_endian PLY in, next to the byte-compatible text formats
    rng = np.random.default_rng(1)
    scr = rng.integers(0, 256, size=(6, 9, 3)).astype(np.float64)
    p6 = tmp_path / "a6.ppm"
    capi.write_ppm_binary(api, p6, scr)
    raw = p6.read_bytes()
    assert raw.startswith(b"P6\n9 6\n255\n") and len(raw) == len(b"P6\n9 6\n255\n") + 6 * 9 * 3
    px = np.frombuffer(raw[len(b"P6\n9 6\n255\n"):], dtype=np.uint8).reshape(6, 9, 3)
    assert np.array_equal(px, scr[::-1].astype(np.uint8))  # rows top-down like the P3 writer
    s = rtb.new_scene()
    assert s.tex_image_ppm(p6) == 0  # P6 accepted as an Image texture source
    p3 = tmp_path / "a3.ppm"
    capi.write_ppm(api, p3, scr)
    assert s.tex_image_ppm(p3) == 1
    # the same mesh as ASCII and as binary PLY flattens to the same scene
    v = rng.uniform(-1, 1, size=(40, 3)).astype(np.float32).astype(np.float64)
    f = rng.integers(0, 40, size=(64, 3)).astype(np.uint32)
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]
    pa, pb, pc = tmp_path / "m_ascii.ply", tmp_path / "m_bin.ply", tmp_path / "m_conv.ply"
    pa.write_text("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement face %d\n"
                  "property list uchar int vertex_indices\nend_header\n" % (len(v), len(f))
                  + "".join("%r %r %r\n" % tuple(float(x) for x in r) for r in v) + "".join("3 %d %d %d\n" % tuple(r) for r in f))
    capi.write_ply_binary(api, pb, v, f)
    capi.ply_convert_binary(api, pa, pc)
    assert pb.read_bytes() == pc.read_bytes()
    checks = []
    for path in (pa, pb):
        g = rtb.new_scene()
        g.set_root(g.bvh([g.ply_load(path, 10.0, g.lambertian((0.2, 0.2, 0.2)))], 0, 1))
        checks.append(g.host_check())
    assert checks[0] == checks[1] and checks[0]["tris"] == len(f) and checks[0]["violations"] == 0
    # a binary PLY with double vertices, an extra vertex property and uint indices
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty double x\nproperty double y\nproperty double z\nproperty uchar q\n"
           "element face 1\nproperty list uchar uint vertex_indices\nend_header\n").encode()
    body = b"".join(np.array(r, dtype="<f8").tobytes() + b"\x07" for r in ((0, 0, 0), (1, 0, 0), (0, 1, 0))) + b"\x03" + np.array([0, 1, 2], dtype="<u4").tobytes()
    pd = tmp_path / "m_d.ply"
    pd.write_bytes(hdr + body)
    g = rtb.new_scene()
    g.set_root(g.bvh([g.ply_load(pd, 1.0, g.lambertian((0.2, 0.2, 0.2)))], 0, 1))
    assert g.host_check()["tris"] == 1
    for bad in (hdr + body[:-2], hdr.replace(b"little", b"big") + body, hdr + body[:-13] + b"\x04" + body[-12:]):
        pd.write_bytes(bad)
        with pytest.raises(capi.RtError):
            g.ply_load(pd, 1.0, 0)


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_commit_fails_loudly_without_gpu(api):
    s = rtb.new_scene()
    s.world_build(4, 1)
    with pytest.raises(capi.RtError) as e:
        s.commit()
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)
    cfg = capi.make_config(32, 1.0, 1, 5)
    with pytest.raises(capi.RtError) as e:
        s.render(cfg)
    assert e.value.code == -2
