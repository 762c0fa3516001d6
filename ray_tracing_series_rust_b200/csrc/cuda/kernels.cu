// kernels.cu — the wavefront path tracer (sm_100a).
//
// Replaces the reference's per-pixel hot loop (src/world.rs:1208-1216 -> Camera::get_ray ->
// ray_color world.rs:52-93 -> Hittable::hit / Material::scatter / Texture::value) with a pool of
// N resident path slots advanced one segment per iteration:
//
//   k_generate   newpath queue -> camera rays (pixel jitter + Camera::get_ray, Philox)      [a1,a2,a23]
//   k_extend     every live slot: world.hit(ray, 0.001, inf) incl. instances and media;
//                miss -> background * throughput into the pixel; hit -> HitRecord into the slot and the
//                slot index into the queue of its material type (warp-aggregated push)      [a3-a15]
//   k_shade<M>   one launch per material type over its queue: scatter / emitted, throughput update,
//                terminated paths add their radiance to the pixel and go to the newpath queue [a16-a22]
//
// Pixels accumulate in int64 fixed point (2^32 scale) with integer atomics: addition is associative,
// so the image is bit-reproducible for a given seed, independent of scheduling, slot count and of
// how the samples are sharded across GPUs.
#include "kernels.h"

#include <algorithm>
#include <cstdio>
#include <vector>

#include "rt_device.cuh"

namespace rtb {

// ------------------------------------------------------------------ path state (SoA, one entry per slot)
struct PathState {
    double *ox, *oy, *oz, *dx, *dy, *dz, *time; // current ray; after a hit (ox,oy,oz) holds HitRecord.p
    double *nx, *ny, *nz;                       // HitRecord.normal
    float *hu, *hv;                             // HitRecord.u, v
    uint32_t* hmat;                             // material id | front_face << 31
    float *tr, *tg, *tb;                        // `product` of ray_color (throughput)
    uint32_t* pixel;                            // j * W + i
    uint64_t* path_id;                          // (j*W + i) * spp_total + sample: the Philox stream id
    uint32_t* draw;                             // Philox draw counter
    uint32_t* segment;                          // ray_color loop iteration
    uint8_t* alive;
};

enum CounterSlot { C_NEWQ0 = 0, C_NEWQ1 = 1, C_MATQ0 = 2, C_DEAD = 7, C_NUM = 8 };
struct Queues {
    uint32_t* newq[2];
    uint32_t* matq[MAT_TYPE_COUNT];
    uint32_t* counts;            // CounterSlot
    unsigned long long* next_path;
    unsigned long long* stats;   // [0] segments, [1] nodes, [2] prims, [3] medium queries, [4..8] scatters by material
};

struct JobDev {
    int32_t W, H, rows, spp_total, sample_begin, max_depth;
    uint32_t npix_rendered;
    unsigned long long total_paths;
    uint64_t seed;
    uint32_t n_slots;
    int32_t count_events;
};

RT_DEV uint32_t lane_id() { return threadIdx.x & 31u; }

// warp-aggregated queue push: lanes of the warp that push to the same counter share one atomic
RT_DEV uint32_t agg_reserve(uint32_t* counter, uint32_t key) {
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, key);
    const int leader = __ffs(grp) - 1;
    const uint32_t rank = __popc(grp & ((1u << lane_id()) - 1u));
    uint32_t base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(grp));
    base = __shfl_sync(grp, base, leader);
    return base + rank;
}

RT_DEV void accumulate(int64_t* __restrict__ accum, uint32_t pixel, float r, float g, float b) {
    // fixed point 2^32 in an int64 (2^31 units of headroom per channel); one sample is clamped to [0, 2^20]
    const double s = 4294967296.0;
    const double lim = 1048576.0;
    double v[3] = {(double)r, (double)g, (double)b};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double x = v[c];
        if (!(x > 0.0)) continue; // 0, negative, NaN contribute nothing
        x = x > lim ? lim : x;
        atomicAdd(reinterpret_cast<unsigned long long*>(accum + (size_t)pixel * 3 + c), (unsigned long long)(long long)(x * s + 0.5));
    }
}

// ------------------------------------------------------------------ k_generate
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ DeviceScene S, const __grid_constant__ JobDev J, PathState P, Queues Q, int cur) {
    const uint32_t n = Q.counts[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int m = 0; m < MAT_TYPE_COUNT; ++m) Q.counts[C_MATQ0 + m] = 0; // consumed by last iteration's shade kernels
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = Q.newq[cur][i];
        // claim the next path index (warp-aggregated)
        const unsigned act = __activemask();
        const int leader = __ffs(act) - 1;
        unsigned long long base = 0;
        if ((int)lane_id() == leader) base = atomicAdd(Q.next_path, (unsigned long long)__popc(act));
        base = __shfl_sync(act, base, leader);
        const unsigned long long L = base + __popc(act & ((1u << lane_id()) - 1u));
        if (L >= J.total_paths) {
            P.alive[slot] = 0;
            atomicAdd(&Q.counts[C_DEAD], 1u);
            continue;
        }
        const uint32_t s_local = (uint32_t)(L / J.npix_rendered);
        const uint32_t pix = (uint32_t)(L % J.npix_rendered);
        const int32_t j = (int32_t)(pix / (uint32_t)J.W), ii = (int32_t)(pix % (uint32_t)J.W);
        const uint64_t path_id = (uint64_t)pix * (uint64_t)J.spp_total + (uint64_t)(J.sample_begin + (int32_t)s_local);
        PathRng g;
        g.init(J.seed, path_id, 0);
        const double u = ((double)ii + g.gen()) / (double)(J.W - 1); // world.rs:1212
        const double v = ((double)j + g.gen()) / (double)(J.H - 1);  // world.rs:1213
        const Ray r = camera_get_ray(S.cam, u, v, g);
        P.ox[slot] = r.o.x; P.oy[slot] = r.o.y; P.oz[slot] = r.o.z;
        P.dx[slot] = r.d.x; P.dy[slot] = r.d.y; P.dz[slot] = r.d.z;
        P.time[slot] = r.time;
        P.tr[slot] = 1.f; P.tg[slot] = 1.f; P.tb[slot] = 1.f;
        P.pixel[slot] = pix;
        P.path_id[slot] = path_id;
        P.draw[slot] = g.draw;
        P.segment[slot] = 0;
        P.alive[slot] = 1;
    }
}

// ------------------------------------------------------------------ k_extend
__global__ void __launch_bounds__(128) k_extend(const __grid_constant__ DeviceScene S, const __grid_constant__ JobDev J, PathState P, Queues Q,
                                                int64_t* __restrict__ accum, int cur) {
    const int nxt = cur ^ 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) Q.counts[cur] = 0; // k_generate has consumed newq[cur]
    uint32_t my_segments = 0;
    TraceCounters tc; tc.nodes = 0; tc.prims = 0;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < J.n_slots; slot += gridDim.x * blockDim.x) {
        if (!P.alive[slot]) continue;
        Ray r;
        r.o = mk3(P.ox[slot], P.oy[slot], P.oz[slot]);
        r.d = mk3(P.dx[slot], P.dy[slot], P.dz[slot]);
        r.time = P.time[slot];
        const uint64_t path_id = S.n_media ? P.path_id[slot] : 0ull;
        const uint32_t segment = S.n_media ? P.segment[slot] : 0u;
        HitRec h;
        bool hit;
        if (J.count_events) hit = world_hit<true, false>(S, r, 0.001, RT_INF, true, J.seed, path_id, segment, h, &tc);
        else hit = world_hit<false, false>(S, r, 0.001, RT_INF, true, J.seed, path_id, segment, h, nullptr);
        ++my_segments;
        if (!hit) {
            // world.rs:86-89: output += product * background; break
            accumulate(accum, P.pixel[slot], P.tr[slot] * S.background[0], P.tg[slot] * S.background[1], P.tb[slot] * S.background[2]);
            Q.newq[nxt][agg_reserve(&Q.counts[nxt], 0u)] = slot;
            continue;
        }
        P.ox[slot] = h.p.x; P.oy[slot] = h.p.y; P.oz[slot] = h.p.z;
        P.nx[slot] = h.n.x; P.ny[slot] = h.n.y; P.nz[slot] = h.n.z;
        P.hu[slot] = (float)h.u; P.hv[slot] = (float)h.v;
        P.hmat[slot] = h.mat | (h.front ? 0x80000000u : 0u);
        const uint32_t mt = __ldg(&S.materials[h.mat].type);
        Q.matq[mt][agg_reserve(&Q.counts[C_MATQ0 + mt], 1u + mt)] = slot;
    }
    // statistics: one atomic per warp
    const unsigned full = 0xffffffffu;
    uint32_t segs = my_segments;
    for (int o = 16; o > 0; o >>= 1) segs += __shfl_xor_sync(full, segs, o);
    if (lane_id() == 0 && segs) atomicAdd(&Q.stats[0], (unsigned long long)segs);
    if (J.count_events) {
        uint32_t a = tc.nodes, b = tc.prims;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(full, a, o); b += __shfl_xor_sync(full, b, o); }
        if (lane_id() == 0) { atomicAdd(&Q.stats[1], (unsigned long long)a); atomicAdd(&Q.stats[2], (unsigned long long)b); }
    }
}

// ------------------------------------------------------------------ k_shade<M>
template <uint32_t M>
__global__ void __launch_bounds__(256) k_shade(const __grid_constant__ DeviceScene S, const __grid_constant__ JobDev J, PathState P, Queues Q,
                                               int64_t* __restrict__ accum, int cur) {
    const int nxt = cur ^ 1;
    const uint32_t n = Q.counts[C_MATQ0 + M];
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&Q.stats[4 + M], (unsigned long long)n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = Q.matq[M][i];
        const uint32_t hm = P.hmat[slot];
        const DMaterial m = S.materials[hm & 0x7fffffffu];
        const D3 p = mk3(P.ox[slot], P.oy[slot], P.oz[slot]);
        bool scattered = false;
        D3 dir = mk3(0, 0, 0);
        F3 att = mkf3(0.f, 0.f, 0.f), emitted = mkf3(0.f, 0.f, 0.f);
        PathRng g;
        if (M == MAT_LIGHT) {
            emitted = tex_value(S, m.tex, (double)P.hu[slot], (double)P.hv[slot], p); // hit.rs:1146-1151
        } else {
            g.init(J.seed, P.path_id[slot], P.draw[slot]);
            const D3 n3 = mk3(P.nx[slot], P.ny[slot], P.nz[slot]);
            if (M == MAT_LAMBERTIAN) scattered = scatter_lambertian(S, m, p, n3, (double)P.hu[slot], (double)P.hv[slot], g, dir, att);
            else if (M == MAT_METAL) scattered = scatter_metal(m, mk3(P.dx[slot], P.dy[slot], P.dz[slot]), n3, g, dir, att);
            else if (M == MAT_DIELECTRIC) scattered = scatter_dielectric(m, mk3(P.dx[slot], P.dy[slot], P.dz[slot]), n3, (hm >> 31) != 0, g, dir, att);
            else scattered = scatter_isotropic(S, m, p, (double)P.hu[slot], (double)P.hv[slot], g, dir, att);
        }
        if (scattered) {
            const uint32_t seg = P.segment[slot] + 1;
            if ((int32_t)seg < J.max_depth) { // world.rs:64-67: at most max_depth hit queries per path
                P.tr[slot] *= att.x; P.tg[slot] *= att.y; P.tb[slot] *= att.z; // world.rs:75
                P.dx[slot] = dir.x; P.dy[slot] = dir.y; P.dz[slot] = dir.z;     // origin is already rec.p
                P.draw[slot] = g.draw;
                P.segment[slot] = seg;
                continue;
            }
            // depth exhausted: the path keeps what it accumulated (nothing, emitters do not scatter)
        } else if (M == MAT_LIGHT) {
            accumulate(accum, P.pixel[slot], P.tr[slot] * emitted.x, P.tg[slot] * emitted.y, P.tb[slot] * emitted.z); // world.rs:78-84
        }
        Q.newq[nxt][agg_reserve(&Q.counts[nxt], 0u)] = slot;
    }
}

// ------------------------------------------------------------------ resolve (vec3.rs:89-107 + Screen layout)
__global__ void k_resolve(const int64_t* __restrict__ accum, double* __restrict__ screen, int32_t W, int32_t H, int32_t spp, int32_t rows) {
    const int64_t n = (int64_t)W * H * 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / 3;
        const int32_t j = (int32_t)(pix / W);
        double out = 0.0;
        if (j < rows) {
            const double sum = (double)accum[i] * (1.0 / 4294967296.0);
            const double scale = 1.0 / (double)spp;
            double c = sqrt(sum * scale);
            c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);     // mutil.rs:1-9 (NaN passes through)
            const double q = 255.9 * c;
            out = (q != q) ? 0.0 : (double)(int32_t)q;    // `as i32`: truncation, NaN -> 0
        }
        screen[i] = out;
    }
}

// ------------------------------------------------------------------ trace_batch (parity hook)
__global__ void __launch_bounds__(128) k_trace_batch(const __grid_constant__ DeviceScene S, const rt_ray* __restrict__ rays, int64_t n, double t_min,
                                                     double t_max, int32_t flags, uint64_t seed, rt_hit* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Ray r;
        r.o = mk3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        r.d = mk3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        r.time = rays[i].time;
        HitRec h;
        const bool hit = world_hit<false, true>(S, r, t_min, t_max, (flags & RT_TRACE_SEEDED_MEDIA) != 0, seed, (uint64_t)i, 0u, h, nullptr);
        rt_hit o;
        if (hit) {
            o.prim_id = (int32_t)h.prim_id; o.mat_id = (int32_t)h.mat; o.t = h.t;
            o.p[0] = h.p.x; o.p[1] = h.p.y; o.p[2] = h.p.z;
            o.normal[0] = h.n.x; o.normal[1] = h.n.y; o.normal[2] = h.n.z;
            o.u = h.u; o.v = h.v; o.front_face = h.front ? 1 : 0; o.pad_ = 0;
        } else {
            o.prim_id = -1; o.mat_id = -1; o.t = 0.0;
            o.p[0] = o.p[1] = o.p[2] = 0.0; o.normal[0] = o.normal[1] = o.normal[2] = 0.0;
            o.u = o.v = 0.0; o.front_face = 0; o.pad_ = 0;
        }
        out[i] = o;
    }
}

// ------------------------------------------------------------------ unit-level device checks
__global__ void k_unit_op(const __grid_constant__ DeviceScene S, int op, uint32_t ia, uint32_t ib, uint32_t ic, uint32_t id, const double* in, double* out) {
    if (threadIdx.x || blockIdx.x) return;
    if (op == 0) {
        const F3 c = tex_value(S, ia, in[0], in[1], mk3(in[2], in[3], in[4]));
        out[0] = c.x; out[1] = c.y; out[2] = c.z;
    } else if (op == 1) {
        out[0] = perlin_noise(&S.perlin[ia], mk3(in[0], in[1], in[2]));
        out[1] = perlin_turbulence(&S.perlin[ia], mk3(in[0], in[1], in[2]), 7);
    } else if (op == 2) {
        const uint4 r = philox4x32_10(make_uint4(ia, ib, ic, id), make_uint2((uint32_t)in[0], (uint32_t)in[1]));
        out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    } else if (op == 3) {
        PathRng g;
        g.init((uint64_t)ia | ((uint64_t)ib << 32), (uint64_t)ic | ((uint64_t)id << 32), 0);
        const double u = (in[0] + g.gen()) / (in[2] - 1.0);
        const double v = (in[1] + g.gen()) / (in[3] - 1.0);
        const Ray r = camera_get_ray(S.cam, u, v, g);
        out[0] = r.o.x; out[1] = r.o.y; out[2] = r.o.z; out[3] = r.d.x; out[4] = r.d.y; out[5] = r.d.z; out[6] = r.time;
    }
}

__global__ void k_iota(uint32_t* q, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) q[i] = i;
}

// ------------------------------------------------------------------ host side
#define CK(x)                                   \
    do {                                        \
        cudaError_t e_ = (x);                   \
        if (e_ != cudaSuccess) { err = e_; goto done; } \
    } while (0)

cudaError_t launch_resolve(const int64_t* d_accum, double* d_screen, int32_t W, int32_t H, int32_t spp, int32_t rows, cudaStream_t stream) {
    const int64_t n = (int64_t)W * H * 3;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    k_resolve<<<blocks, 256, 0, stream>>>(d_accum, d_screen, W, H, spp, rows);
    return cudaGetLastError();
}

cudaError_t launch_trace_batch(const DeviceScene& scene, const rt_ray* d_rays, int64_t n, double t_min, double t_max, int32_t flags, uint64_t seed,
                               rt_hit* d_out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int blocks = (int)std::min<int64_t>((n + 127) / 128, 148 * 32);
    k_trace_batch<<<blocks, 128, 0, stream>>>(scene, d_rays, n, t_min, t_max, flags, seed, d_out);
    return cudaGetLastError();
}

cudaError_t launch_unit_op(const DeviceScene& scene, int op, uint32_t ia, uint32_t ib, uint32_t ic, uint32_t id, const double* in8, double* out8) {
    double *d_in = nullptr, *d_out = nullptr;
    cudaError_t err = cudaSuccess;
    CK(cudaMalloc(&d_in, 8 * sizeof(double)));
    CK(cudaMalloc(&d_out, 8 * sizeof(double)));
    CK(cudaMemcpy(d_in, in8, 8 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0, 8 * sizeof(double)));
    k_unit_op<<<1, 32>>>(scene, op, ia, ib, ic, id, d_in, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out8, d_out, 8 * sizeof(double), cudaMemcpyDeviceToHost));
done:
    cudaFree(d_in);
    cudaFree(d_out);
    return err;
}

namespace {
struct Arena { // one allocation for the whole path state + queues
    char* base = nullptr;
    size_t used = 0, cap = 0;
    template <class T> T* take(size_t n) {
        used = (used + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
};
} // namespace

cudaError_t launch_render(const DeviceScene& scene, const RenderJob& job, const RenderTuning& tune, int64_t* d_accum, cudaStream_t stream,
                          rt_stats* stats) {
    cudaError_t err = cudaSuccess;
    JobDev J;
    J.W = job.width; J.H = job.height; J.rows = job.rows; J.spp_total = job.spp_total;
    J.sample_begin = job.sample_begin; J.max_depth = job.max_depth;
    J.npix_rendered = (uint32_t)job.width * (uint32_t)job.rows;
    J.total_paths = (unsigned long long)J.npix_rendered * (unsigned long long)(job.sample_end - job.sample_begin);
    J.seed = job.seed;
    J.count_events = tune.count_events;
    uint32_t N = tune.wave_slots;
    if ((unsigned long long)N > J.total_paths) N = (uint32_t)std::max<unsigned long long>(J.total_paths, 1ull);
    N = (N + 127u) & ~127u;
    J.n_slots = N;

    Arena A;
    PathState P;
    Queues Q;
    unsigned int* h_flag = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_poll[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ext_events;
    unsigned long long h_stats[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t iterations = 0, launches = 0;
    float ms_device = 0.f;
    double ms_extend = 0.0;

    if (J.total_paths == 0) goto done;
    {
        const size_t per_slot = 7 * 8 + 3 * 8 + 2 * 4 + 4 + 3 * 4 + 4 + 8 + 4 + 4 + 1 + (2 + MAT_TYPE_COUNT) * 4;
        A.cap = (size_t)N * per_slot + 64 * 256 + 4096;
        CK(cudaMalloc(&A.base, A.cap));
        P.ox = A.take<double>(N); P.oy = A.take<double>(N); P.oz = A.take<double>(N);
        P.dx = A.take<double>(N); P.dy = A.take<double>(N); P.dz = A.take<double>(N); P.time = A.take<double>(N);
        P.nx = A.take<double>(N); P.ny = A.take<double>(N); P.nz = A.take<double>(N);
        P.hu = A.take<float>(N); P.hv = A.take<float>(N); P.hmat = A.take<uint32_t>(N);
        P.tr = A.take<float>(N); P.tg = A.take<float>(N); P.tb = A.take<float>(N);
        P.pixel = A.take<uint32_t>(N); P.path_id = A.take<uint64_t>(N); P.draw = A.take<uint32_t>(N); P.segment = A.take<uint32_t>(N);
        P.alive = A.take<uint8_t>(N);
        Q.newq[0] = A.take<uint32_t>(N); Q.newq[1] = A.take<uint32_t>(N);
        for (int m = 0; m < MAT_TYPE_COUNT; ++m) Q.matq[m] = A.take<uint32_t>(N);
        Q.counts = A.take<uint32_t>(C_NUM);
        Q.next_path = A.take<unsigned long long>(1);
        Q.stats = A.take<unsigned long long>(9);
        if (A.used > A.cap) { err = cudaErrorMemoryAllocation; goto done; }
    }
    CK(cudaMallocHost(&h_flag, 2 * sizeof(unsigned int)));
    CK(cudaEventCreate(&ev_begin));
    CK(cudaEventCreate(&ev_end));
    CK(cudaEventCreateWithFlags(&ev_poll[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_poll[1], cudaEventDisableTiming));
    {
        CK(cudaMemsetAsync(P.alive, 0, N, stream));
        CK(cudaMemsetAsync(Q.counts, 0, C_NUM * sizeof(uint32_t), stream));
        CK(cudaMemsetAsync(Q.next_path, 0, sizeof(unsigned long long), stream));
        CK(cudaMemsetAsync(Q.stats, 0, 9 * sizeof(unsigned long long), stream));
        const int qblocks = (int)std::min<uint32_t>((N + 255) / 256, 148 * 8);
        k_iota<<<qblocks, 256, 0, stream>>>(Q.newq[0], N);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(Q.counts + C_NEWQ0, &N, sizeof(uint32_t), cudaMemcpyHostToDevice, stream)); // pageable source: copied before return

        const int eblocks = (int)std::min<uint32_t>((N + 127) / 128, 148 * 16);
        CK(cudaEventRecord(ev_begin, stream));
        int cur = 0;
        const int batch = 16; // iterations enqueued between completion checks
        int pending = -1;     // index of the poll slot still in flight
        bool finished = false;
        h_flag[0] = h_flag[1] = 0;
        for (uint64_t guard = 0; !finished && guard < (1ull << 40); ++guard) {
            const int slot = (int)(guard & 1);
            for (int it = 0; it < batch; ++it) {
                k_generate<<<qblocks, 256, 0, stream>>>(scene, J, P, Q, cur);
                if (tune.timed_extend) {
                    cudaEvent_t a, b;
                    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
                    CK(cudaEventRecord(a, stream));
                    k_extend<<<eblocks, 128, 0, stream>>>(scene, J, P, Q, d_accum, cur);
                    CK(cudaEventRecord(b, stream));
                    ext_events.push_back(a); ext_events.push_back(b);
                } else {
                    k_extend<<<eblocks, 128, 0, stream>>>(scene, J, P, Q, d_accum, cur);
                }
                k_shade<MAT_LAMBERTIAN><<<qblocks, 256, 0, stream>>>(scene, J, P, Q, d_accum, cur);
                k_shade<MAT_METAL><<<qblocks, 256, 0, stream>>>(scene, J, P, Q, d_accum, cur);
                k_shade<MAT_DIELECTRIC><<<qblocks, 256, 0, stream>>>(scene, J, P, Q, d_accum, cur);
                k_shade<MAT_LIGHT><<<qblocks, 256, 0, stream>>>(scene, J, P, Q, d_accum, cur);
                k_shade<MAT_ISOTROPIC><<<qblocks, 256, 0, stream>>>(scene, J, P, Q, d_accum, cur);
                cur ^= 1;
                ++iterations;
                launches += 7;
            }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&h_flag[slot], Q.counts + C_DEAD, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
            CK(cudaEventRecord(ev_poll[slot], stream));
            if (pending >= 0) {
                CK(cudaEventSynchronize(ev_poll[pending]));
                if (h_flag[pending] >= N) finished = true;
            }
            pending = slot;
        }
        if (!finished) {
            // drain: the last enqueued batch may have finished the job
            CK(cudaEventSynchronize(ev_poll[pending]));
        }
        CK(cudaEventRecord(ev_end, stream));
        CK(cudaMemcpyAsync(h_stats, Q.stats, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaEventElapsedTime(&ms_device, ev_begin, ev_end));
        for (size_t i = 0; i + 1 < ext_events.size(); i += 2) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ext_events[i], ext_events[i + 1]));
            ms_extend += ms;
        }
    }
done:
    if (stats) {
        stats->paths = J.total_paths;
        stats->segments = h_stats[0];
        stats->box_tests = h_stats[1];
        stats->prim_tests[0] = h_stats[2];
        for (int m = 0; m < 5; ++m) stats->scatters[m] = h_stats[4 + m];
        stats->iterations = iterations;
        stats->kernel_launches = launches;
        stats->ms_device = ms_device;
        stats->ms_extend = ms_extend;
    }
    for (cudaEvent_t e : ext_events) cudaEventDestroy(e);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (ev_poll[0]) cudaEventDestroy(ev_poll[0]);
    if (ev_poll[1]) cudaEventDestroy(ev_poll[1]);
    if (h_flag) cudaFreeHost(h_flag);
    if (A.base) cudaFree(A.base);
    return err;
}

} // namespace rtb
