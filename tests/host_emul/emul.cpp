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

extern "C" {

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

// stats[0] = wide nodes visited, [1] = leaves tested, [2] = deepest stack (only meaningful for resume_rounds == 0)
int32_t emul_trace_wide(const void* scene, const rt_ray* rays, int64_t n, double t_min, double t_max, int32_t resume, rt_hit* out) {
    const DeviceScene& S = *static_cast<const DeviceScene*>(scene);
    if (!S.nodes4) return -1;
    for (int64_t i = 0; i < n; ++i) {
        const Ray r = load_ray(rays[i]);
        BestHit best;
        best_init(best, t_max);
        unsigned long long stack[RT_WIDE_STACK];
        uint32_t cur = S.root4;
        int sp = 0;
        if (resume) { // one lane: every call returns after the lane finishes; wait_thresh 1 exercises the state hand-over
            while (cur != 0xffffffffu) trace_wide<RT_PM_ALL, true>(S, r, t_min, best, cur, sp, stack, 1u);
        } else {
            trace_wide<RT_PM_ALL, false>(S, r, t_min, best, cur, sp, stack, 0u);
        }
        HitRec h;
        const bool hit = best.type != RT_NONE;
        if (hit) h = finalize_hit<1, RT_PM_ALL, false>(S, r, best);
        store_hit(out[i], hit, h);
    }
    return 0;
}

} // extern "C"
