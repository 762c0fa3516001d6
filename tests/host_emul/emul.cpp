// emul.cpp — TEST INFRASTRUCTURE (tests/test_host_emul.py).  The device header compiled for the host:
//   emul_trace_batch   = k_trace_batch's body (world.hit through the sibling-pair walk, media included)
//   emul_trace_wide    = closest hit through trace_wide (the 4-wide walk of k_mega_r), resumable or not
// both over the host-flattened DeviceScene that rt_debug_host_scene hands out.  -ffp-contract=off: every FMA in
// rt_device.cuh is written out, as in the -fmad=false device build.
#include "cuda_shim.h"

#include "../../include/rtb200.h"
#include "../../ray_tracing_series_rust_b200/csrc/cuda/rt_device.cuh"

using namespace rtb;

static void store_hit(rt_hit& o, bool hit, const HitRec& h) {
    std::memset(&o, 0, sizeof o);
    if (hit) {
        o.prim_id = (int32_t)h.prim_id; o.mat_id = (int32_t)h.mat; o.t = h.t;
        o.p[0] = h.p.x; o.p[1] = h.p.y; o.p[2] = h.p.z;
        o.normal[0] = h.n.x; o.normal[1] = h.n.y; o.normal[2] = h.n.z;
        o.u = h.u; o.v = h.v; o.front_face = h.front ? 1 : 0;
    } else {
        o.prim_id = -1; o.mat_id = -1;
    }
}
static Ray load_ray(const rt_ray& q) {
    Ray r;
    r.o = mk3(q.o[0], q.o[1], q.o[2]);
    r.d = mk3(q.d[0], q.d[1], q.d[2]);
    r.time = q.time;
    return r;
}

// Closest hit through trace_wide as the fused kernels call it (single main instance).  pm3 != 0: the spheres + moving-spheres
// instantiation, which walks the motion form of the wide nodes (DeviceScene::mnodes4) when the scene has it.
template <uint32_t PM>
static void wide_batch(const DeviceScene& S, const rt_ray* rays, int64_t n, double t_min, double t_max, int32_t resume, rt_hit* out) {
    for (int64_t i = 0; i < n; ++i) {
        const Ray r = load_ray(rays[i]);
        BestHit best;
        best_init(best, t_max);
        unsigned long long stack[RT_WIDE_STACK];
        uint32_t cur = S.root4;
        int sp = 0;
        if (resume) { // one lane: every call returns after the lane finishes; wait_thresh 1 exercises the state hand-over
            while (cur != 0xffffffffu) trace_wide<PM, true>(S, r, t_min, best, cur, sp, stack, 1u);
        } else {
            trace_wide<PM, false>(S, r, t_min, best, cur, sp, stack, 0u);
        }
        HitRec h;
        const bool hit = best.type != RT_NONE;
        if (hit) h = finalize_hit<1, PM, false>(S, r, best);
        store_hit(out[i], hit, h);
    }
}
// A stack that counts its accesses per index: how deep the 4-wide walk's stack really gets (sizes WStackSm's shared-memory part)
namespace rtb {
struct WStackStat {
    unsigned long long e[RT_WIDE_STACK];
    uint64_t* hist; // [RT_WIDE_STACK] puts + gets per index
};
inline void wstk_put(WStackStat& s, int i, unsigned long long v) { s.e[i] = v; ++s.hist[i]; }
inline unsigned long long wstk_get(const WStackStat& s, int i) { ++s.hist[i]; return s.e[i]; }
} // namespace rtb

extern "C" {

// Paths of the fused WIDE kernels (single main instance, no wrappers, no media) with the counting stack; hist[RT_WIDE_STACK].
int32_t emul_wide_stack_hist(const void* scene, int32_t W, int32_t H, int32_t spp, int32_t max_depth, uint64_t seed, uint64_t* hist, uint64_t* segments_out) {
    const DeviceScene& S = *static_cast<const DeviceScene*>(scene);
    if (!S.nodes4) return -1;
    uint64_t segments = 0;
    for (int32_t j = 0; j < H; ++j)
        for (int32_t ii = 0; ii < W; ++ii)
            for (int32_t sl = 0; sl < spp; ++sl) {
                const uint32_t pixel = (uint32_t)j * (uint32_t)W + (uint32_t)ii;
                const uint64_t path_id = (uint64_t)pixel * (uint64_t)spp + (uint64_t)sl;
                uint32_t draw = 0;
                Ray r = camera_first_ray<PathRng>(S.cam, ii, j, W, H, seed, path_id, draw);
                for (uint32_t segment = 0;; ++segment) {
                    BestHit best;
                    best_init(best, RT_INF);
                    WStackStat st;
                    st.hist = hist;
                    uint32_t cur = S.root4;
                    int sp = 0;
                    trace_wide<RT_PM_ALL, false>(S, r, 0.001, best, cur, sp, st, 0u);
                    ++segments;
                    if (best.type == RT_NONE) break;
                    const HitRec h = finalize_hit<2, RT_PM_ALL, false>(S, r, best);
                    const DMaterial m = S.materials[h.mat];
                    if (m.type == MAT_LIGHT) break;
                    PathRng g;
                    g.init(seed, path_id, draw);
                    D3 dir = mk3(0, 0, 0);
                    F3 att = mkf3(0.f, 0.f, 0.f);
                    bool scattered;
                    if (m.type == MAT_LAMBERTIAN) scattered = scatter_lambertian(S, m, h.p, h.n, h.u, h.v, g, dir, att);
                    else if (m.type == MAT_METAL) scattered = scatter_metal(m, r.d, h.n, g, dir, att);
                    else if (m.type == MAT_DIELECTRIC) scattered = scatter_dielectric(m, r.d, h.n, h.front, g, dir, att);
                    else scattered = scatter_isotropic(S, m, h.p, h.u, h.v, g, dir, att);
                    if (!(scattered && (int32_t)(segment + 1) < max_depth)) break;
                    r.o = h.p; r.d = dir;
                    draw = g.draw;
                }
            }
    if (segments_out) *segments_out = segments;
    return 0;
}

// PathRng's conversions against the formulas they replace ((double)u * 2^-32 and fma(xi, 2, -1)); -> number of mismatching u
uint64_t emul_rng_convert_mismatches(uint64_t n_random, uint64_t seed) {
    uint64_t bad = 0, x = seed | 1u;
    auto check = [&](uint32_t u) {
        const double xi = (double)u * (1.0 / 4294967296.0);
        const double a = PathRng::unit_of(u), b = PathRng::pm1_of(u), b0 = fma(xi, 1.0 - -1.0, -1.0);
        if (std::memcmp(&a, &xi, 8) != 0 || std::memcmp(&b, &b0, 8) != 0) ++bad;
    };
    const uint32_t edge[] = {0u, 1u, 2u, 0x7fffffffu, 0x80000000u, 0x80000001u, 0xfffffffeu, 0xffffffffu, 0x00100000u, 0x001fffffu};
    for (uint32_t u : edge) check(u);
    for (uint64_t i = 0; i < n_random; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; check((uint32_t)(x >> 16)); }
    return bad;
}

// next3_pm1 / next2_pm1 from every draw position 0..23 (aligned or not, block cached or not), interleaved with single draws, against a
// second generator that only ever calls gen_range(-1, 1) / gen(): same values, same draw counter afterwards; -> number of mismatches
uint64_t emul_rng_multi_draw_mismatches(uint64_t seed, uint64_t n_paths) {
    uint64_t bad = 0;
    auto same = [&](double x, double y) { if (std::memcmp(&x, &y, 8) != 0) ++bad; };
    for (uint64_t path = 0; path < n_paths; ++path)
        for (uint32_t start = 0; start < 24; ++start)
            for (int warm = 0; warm < 2; ++warm) { // warm = 1: the block of `start` is already cached (one gen() taken first, from start - 1)
                PathRng a, b;
                const uint32_t s0 = warm && start ? start - 1 : start;
                a.init(seed, path, s0); b.init(seed, path, s0);
                if (warm && start) same(a.gen(), b.gen());
                for (int round = 0; round < 5; ++round) {
                    double x, y, z;
                    if ((round + start) & 1) { a.next3_pm1(x, y, z); same(x, b.gen_range(-1.0, 1.0)); same(y, b.gen_range(-1.0, 1.0)); same(z, b.gen_range(-1.0, 1.0)); }
                    else { a.next2_pm1(x, y); same(x, b.gen_range(-1.0, 1.0)); same(y, b.gen_range(-1.0, 1.0)); }
                    if (round == 2) same(a.gen(), b.gen());
                    if (a.draw != b.draw) ++bad;
                }
                PathRngOolBegun c; // the fused kernels' generator: open_event == begin_event of the plain one, begin_event a no-op
                PathRng d;
                c.init(seed, path, start); d.init(seed, path, start);
                c.open_event(); c.begin_event(); d.begin_event();
                double x, y, z;
                c.next3_pm1(x, y, z);
                same(x, d.gen_range(-1.0, 1.0)); same(y, d.gen_range(-1.0, 1.0)); same(z, d.gen_range(-1.0, 1.0));
                if (c.draw != d.draw) ++bad;
            }
    return bad;
}

uint64_t emul_sizeof_device_scene(void) { return sizeof(DeviceScene); }

// wide != 0: the main world through Instance::root4 (world_hit<..., WIDE = true>, what k_extend<..., WIDE> runs); needs t_min >= 0.
// counts (may be null): [0] boxes tested, [1] primitives tested, summed over the batch
int32_t emul_trace_batch(const void* scene, const rt_ray* rays, int64_t n, double t_min, double t_max, int32_t flags, uint64_t seed, int32_t wide,
                         rt_hit* out, uint64_t* counts) {
    const DeviceScene& S = *static_cast<const DeviceScene*>(scene);
    if (wide && (!S.nodes4 || !(t_min >= 0.0))) return -1;
    uint64_t nodes = 0, prims = 0;
    for (int64_t i = 0; i < n; ++i) {
        HitRec h;
        TraceCounters tc; tc.nodes = 0; tc.prims = 0;
        const bool media = (flags & RT_TRACE_SEEDED_MEDIA) != 0;
        const bool hit = wide ? world_hit<true, 1, true, true, RT_PM_ALL, true, true>(S, load_ray(rays[i]), t_min, t_max, media, seed, (uint64_t)i, 0u, h, &tc)
                              : world_hit<true, 1, true>(S, load_ray(rays[i]), t_min, t_max, media, seed, (uint64_t)i, 0u, h, &tc);
        nodes += tc.nodes; prims += tc.prims;
        store_hit(out[i], hit, h);
    }
    if (counts) { counts[0] = nodes; counts[1] = prims; }
    return 0;
}

int32_t emul_trace_wide(const void* scene, const rt_ray* rays, int64_t n, double t_min, double t_max, int32_t resume, int32_t pm3, rt_hit* out) {
    const DeviceScene& S = *static_cast<const DeviceScene*>(scene);
    if (!S.nodes4) return -1;
    if (pm3) {
        if (!S.mnodes4) return -2;
        wide_batch<0x3u>(S, rays, n, t_min, t_max, resume, out);
    } else {
        wide_batch<RT_PM_ALL>(S, rays, n, t_min, t_max, resume, out);
    }
    return 0;
}

// One path of k_mega / k_shade_all + k_extend, lane by lane: camera_first_ray -> (world_hit -> emitted / scatter)* -> fixed-point
// accumulate (kernels.cu accumulate()).  Same device functions, same Philox streams as the kernels; wide bit 0 walks Instance::root4,
// bit 1 shades in the fused kernels' order of operations.
int32_t emul_render(const void* scene, int32_t W, int32_t H, int32_t spp_total, int32_t sample_begin, int32_t sample_end, int32_t max_depth, uint64_t seed,
                    int32_t wide, int64_t* accum, uint64_t* segments_out) {
    const DeviceScene& S = *static_cast<const DeviceScene*>(scene);
    const bool fused_style = (wide & 2) != 0; // bit 1: shade as the fused kernels do (same draws in the same order, restructured)
    wide &= 1;
    if (wide && !S.nodes4) return -1;
    uint64_t segments = 0;
    for (int32_t j = 0; j < H; ++j)
        for (int32_t ii = 0; ii < W; ++ii)
            for (int32_t sl = sample_begin; sl < sample_end; ++sl) {
                const uint32_t pixel = (uint32_t)j * (uint32_t)W + (uint32_t)ii;
                const uint64_t path_id = (uint64_t)pixel * (uint64_t)spp_total + (uint64_t)sl;
                uint32_t draw = 0;
                Ray r = camera_first_ray<PathRng>(S.cam, ii, j, W, H, seed, path_id, draw);
                float tr = 1.f, tg = 1.f, tb = 1.f;
                F3 contrib = mkf3(0.f, 0.f, 0.f);
                for (uint32_t segment = 0;; ++segment) {
                    HitRec h;
                    const bool hit = wide ? world_hit<false, 2, true, true, RT_PM_ALL, true, true>(S, r, 0.001, RT_INF, true, seed, path_id, segment, h, nullptr)
                                          : world_hit<false, 2, true>(S, r, 0.001, RT_INF, true, seed, path_id, segment, h, nullptr);
                    ++segments;
                    if (!hit) { const F3 bg = miss_color(S, r.d); contrib = mkf3(tr * bg.x, tg * bg.y, tb * bg.z); break; }
                    const DMaterial m = S.materials[h.mat];
                    if (m.type == MAT_LIGHT) {
                        const F3 e = tex_value(S, m.tex, h.u, h.v, h.p);
                        contrib = mkf3(tr * e.x, tg * e.y, tb * e.z);
                        break;
                    }
                    D3 dir = mk3(0, 0, 0);
                    F3 att = mkf3(0.f, 0.f, 0.f);
                    bool scattered;
                    if (fused_style) { // as k_mega / k_mega_r shade: event opened before the material branch, one rejection loop for three materials
                        PathRngOolBegun g;
                        g.init(seed, path_id, draw);
                        g.open_event();
                        D3 rs = mk3(0, 0, 0);
                        if (m.type != MAT_DIELECTRIC) rs = random_in_unit_sphere(g);
                        if (m.type == MAT_LAMBERTIAN) scattered = lambertian_finish(S, m, h.p, h.n, h.u, h.v, rs, dir, att);
                        else if (m.type == MAT_METAL) scattered = metal_finish(m, r.d, h.n, rs, dir, att);
                        else if (m.type == MAT_DIELECTRIC) scattered = scatter_dielectric(m, r.d, h.n, h.front, g, dir, att);
                        else scattered = isotropic_finish(S, m, h.p, h.u, h.v, rs, dir, att);
                        draw = g.draw;
                    } else { // as k_shade_all shades
                        PathRng g;
                        g.init(seed, path_id, draw);
                        if (m.type == MAT_LAMBERTIAN) scattered = scatter_lambertian(S, m, h.p, h.n, h.u, h.v, g, dir, att);
                        else if (m.type == MAT_METAL) scattered = scatter_metal(m, r.d, h.n, g, dir, att);
                        else if (m.type == MAT_DIELECTRIC) scattered = scatter_dielectric(m, r.d, h.n, h.front, g, dir, att);
                        else scattered = scatter_isotropic(S, m, h.p, h.u, h.v, g, dir, att);
                        draw = g.draw;
                    }
                    if (!(scattered && (int32_t)(segment + 1) < max_depth)) break;
                    tr *= att.x; tg *= att.y; tb *= att.z;
                    r.o = h.p; r.d = dir;
                }
                const double v[3] = {(double)contrib.x, (double)contrib.y, (double)contrib.z};
                for (int c = 0; c < 3; ++c) { // kernels.cu accumulate(): 2^32 fixed point, one sample clamped to [0, 2^20]
                    double x = v[c];
                    if (!(x > 0.0)) continue;
                    x = x > 1048576.0 ? 1048576.0 : x;
                    accum[(size_t)pixel * 3 + c] += (int64_t)(x * 4294967296.0 + 0.5);
                }
            }
    if (segments_out) *segments_out = segments;
    return 0;
}

} // extern "C"
