#!/usr/bin/env python
"""Can the primitive tests run in f32 and still meet the north-star bar (ids identical except grazing cases, t within 1e-5 relative)?

A numerical A/B on the book-1 final geometry (CPU, numpy; no GPU needed): the reference's Sphere::hit (hit.rs:204-238) in f64 is the
truth; against it, for the same rays, (a) the same formula evaluated in f32 and (b) a cancellation-free f32 form (Haines et al., "Precision
improvements for ray / sphere intersection", Ray Tracing Gems ch. 7: discriminant from the distance of the centre to the ray, root from
the numerically stable quadratic), both with the ray itself rounded to f32, as a kernel that keeps its rays in f32 registers would see it.
Rays: camera rays of the book-1 camera and secondary rays leaving the hit points in cosine-like directions (t_min = 0.001).

    python tools/f32_study.py            -> prints the table kept in DESIGN.md section 3
"""
import numpy as np

rng = np.random.default_rng(0xB001)


def scene():
    c, r = [(0.0, -1000.0, 0.0)], [1000.0]
    for a in range(-11, 11):
        for b in range(-11, 11):
            p = (a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
            if np.hypot(p[0] - 4.0, p[2]) > 0.9:
                c.append(p); r.append(0.2)
    for p in ((0, 1, 0), (-4, 1, 0), (4, 1, 0)):
        c.append(p); r.append(1.0)
    return np.array(c), np.array(r)


def hit_ref(o, d, c, r, t_min, dtype):
    """hit.rs:204-222 for every (ray, sphere) pair in `dtype`; returns (t, id) of the closest hit (id -1 = miss)"""
    o, d, c, r = (x.astype(dtype) for x in (o, d, c, r))
    best_t = np.full(o.shape[0], np.inf, dtype=dtype)
    best_i = np.full(o.shape[0], -1)
    a = (d * d).sum(1)
    for i in range(c.shape[0]):
        oc = o - c[i]
        hb = (oc * d).sum(1)
        cc = (oc * oc).sum(1) - r[i] * r[i]
        disc = hb * hb - a * cc
        ok = disc >= 0
        sq = np.sqrt(np.where(ok, disc, 0))
        t0, t1 = (-hb - sq) / a, (-hb + sq) / a
        t = np.where((t0 >= t_min) & (t0 <= best_t), t0, t1)
        ok &= (t >= t_min) & (t <= best_t)
        best_t = np.where(ok, t, best_t)
        best_i = np.where(ok, i, best_i)
    return best_t.astype(np.float64), best_i


def hit_robust_f32(o, d, c, r, t_min):
    """the cancellation-free form in f32"""
    f = np.float32
    o, d, c, r = (x.astype(f) for x in (o, d, c, r))
    best_t = np.full(o.shape[0], np.inf, dtype=f)
    best_i = np.full(o.shape[0], -1)
    a = (d * d).sum(1)
    inv_len = f(1) / np.sqrt(a)
    dn = d * inv_len[:, None]
    for i in range(c.shape[0]):
        oc = o - c[i]                      # f = o - c
        b = -(oc * dn).sum(1)              # b' = -f.d^
        l = oc + b[:, None] * dn           # f + b' d^ : centre-to-ray offset
        disc = r[i] * r[i] - (l * l).sum(1)
        ok = disc >= 0
        sq = np.sqrt(np.where(ok, disc, 0))
        q = b + np.where(b >= 0, sq, -sq)  # stable root pair
        cc = (oc * oc).sum(1) - r[i] * r[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            ta, tb = cc / q, q
        t0, t1 = np.minimum(ta, tb) * inv_len, np.maximum(ta, tb) * inv_len
        t = np.where((t0 >= t_min) & (t0 <= best_t), t0, t1)
        ok &= (t >= t_min) & (t <= best_t)
        best_t = np.where(ok, t, best_t)
        best_i = np.where(ok, i, best_i)
    return best_t.astype(np.float64), best_i


def report(name, o, d, c, r, t_min):
    t64, i64 = hit_ref(o, d, c, r, t_min, np.float64)
    rows = []
    for label, (t, i) in (("same formula, f32", hit_ref(o, d, c, r, t_min, np.float32)), ("cancellation-free, f32", hit_robust_f32(o, d, c, r, t_min))):
        hit = i64 >= 0
        same = hit & (i == i64)
        rel = np.abs(t[same] - t64[same]) / np.abs(t64[same])
        ground = i64[same] == 0
        rows.append((label, float((i != i64).mean()), float((rel > 1e-5).mean()), float((rel[ground] > 1e-5).mean()) if ground.any() else 0.0,
                     float((rel[~ground] > 1e-5).mean()) if (~ground).any() else 0.0, float(np.median(rel[ground])) if ground.any() else 0.0))
    print(f"{name}: {o.shape[0]} rays, {int((i64 >= 0).sum())} hits ({int((i64 == 0).sum())} on the r = 1000 ground sphere)")
    for row in rows:
        print("   %-24s id mismatch %.4f | t off by > 1e-5 rel: all %.4f, ground sphere %.4f, small spheres %.5f | median rel err on the ground %.1e" % row)
    return t64, i64


def main():
    c, r = scene()
    n = 60000
    # book-1 camera (world.rs:1157-1177 preset, camera.rs:20-57): lookfrom (13,2,3) -> 0, vfov 20, aspect 3/2, focus 10
    lookfrom = np.array([13.0, 2.0, 3.0]); w = lookfrom / np.linalg.norm(lookfrom)
    u = np.cross([0, 1, 0], w); u /= np.linalg.norm(u); v = np.cross(w, u)
    h = np.tan(np.radians(20) / 2); vh, vw = 2 * h, 1.5 * 2 * h
    s, t = rng.random(n), rng.random(n)
    d = (lookfrom - 10 * vw / 2 * u - 10 * vh / 2 * v - 10 * w) + s[:, None] * (10 * vw * u) + t[:, None] * (10 * vh * v) - lookfrom
    o = np.tile(lookfrom, (n, 1))
    t64, i64 = report("camera rays", o, d, c, r, 0.001)
    hit = i64 >= 0
    p = o[hit] + t64[hit, None] * d[hit]
    nrm = (p - c[i64[hit]]) / r[i64[hit], None]
    rd = rng.normal(size=p.shape); rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    report("secondary rays (origin on a surface, t_min 0.001)", p, nrm + rd, c, r, 0.001)


if __name__ == "__main__":
    main()
