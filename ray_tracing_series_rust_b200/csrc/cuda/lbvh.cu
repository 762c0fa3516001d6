// lbvh.cu — device-side BVH build, sm_100a.
//
// Replaces BvhNode::new (src/bvh.rs:14-83: recursive median split over cloned object vectors, Θ(N²) clones) for
// large primitive sets with a linear BVH built on the GPU:
//
//   k_lbvh_keys     centroid -> 20 bit/axis Morton code, key = type << 60 | morton   (type-major keys keep
//                   every leaf homogeneous and make typed indices = sorted position - type offset)
//   cub radix sort  (key, prim index) pairs, 63 significant bits
//   k_lbvh_topology Karras 2012: internal node i covers a contiguous key range; children from the split
//                   (ties between equal keys are broken by the sorted index)
//   k_lbvh_refit    bottom-up boxes with one atomic ticket per internal node
//   k_lbvh_mark     subtrees of <= max_leaf primitives collapse into one leaf; exclusive scan numbers the
//                   internal nodes that stay
//   k_lbvh_emit     writes the 32-byte sibling-pair layout of rt_types.h that the traversal kernels walk
//
// The closest hit does not depend on tree topology (bvh.rs:97-112), so any tree over the same outward-rounded
// boxes is a legal replacement; tests/test_gpu_parity.py checks hits against the oracle for both builders.
// The tree is of lower quality than the host's binned-SAH tree (more boxes per ray), so it is opt-in:
// rt_scene_set_bvh_builder(RT_BVH_DEVICE_LBVH) / RTB200_BVH_BUILDER=lbvh.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdint>
#include <vector>

#include "kernels.h"

namespace rtb {

namespace {

struct Box6 {
    float lo[3], hi[3];
};

__device__ __forceinline__ uint64_t spread20(uint32_t v) { // 20 bits -> every third bit of 60
    uint64_t x = v & 0xfffffu;
    x = (x | (x << 32)) & 0x001f00000000ffffull;
    x = (x | (x << 16)) & 0x001f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

__global__ void k_lbvh_keys(const Box6* __restrict__ boxes, const uint8_t* __restrict__ types, uint32_t n, float3 cmin, float3 cinv, uint64_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Box6 b = boxes[i];
    const float cx = 0.5f * (b.lo[0] + b.hi[0]), cy = 0.5f * (b.lo[1] + b.hi[1]), cz = 0.5f * (b.lo[2] + b.hi[2]);
    const float fx = fminf(fmaxf((cx - cmin.x) * cinv.x, 0.f), 1.f), fy = fminf(fmaxf((cy - cmin.y) * cinv.y, 0.f), 1.f),
                fz = fminf(fmaxf((cz - cmin.z) * cinv.z, 0.f), 1.f);
    const uint32_t qx = min((uint32_t)(fx * 1048576.f), 1048575u), qy = min((uint32_t)(fy * 1048576.f), 1048575u), qz = min((uint32_t)(fz * 1048576.f), 1048575u);
    keys[i] = ((uint64_t)types[i] << 60) | (spread20(qx) << 2) | (spread20(qy) << 1) | spread20(qz);
    vals[i] = i;
}

// Length of the common prefix of keys i and j (Karras 2012, section 4); equal keys fall back to the index.
__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

#define LBVH_LEAF 0x80000000u

// One thread per internal node i in [0, n-1): its key range and its two children.
__global__ void k_lbvh_topology(const uint64_t* __restrict__ keys, int n, uint32_t* __restrict__ left, uint32_t* __restrict__ right, uint32_t* __restrict__ first,
                                uint32_t* __restrict__ last, uint32_t* __restrict__ parent_int, uint32_t* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t lc = (lo == gamma) ? (LBVH_LEAF | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t rc = (hi == gamma + 1) ? (LBVH_LEAF | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    left[i] = lc; right[i] = rc;
    first[i] = (uint32_t)lo; last[i] = (uint32_t)hi;
    if (lc & LBVH_LEAF) parent_leaf[gamma] = (uint32_t)i; else parent_int[gamma] = (uint32_t)i;
    if (rc & LBVH_LEAF) parent_leaf[gamma + 1] = (uint32_t)i; else parent_int[gamma + 1] = (uint32_t)i;
    if (i == 0) parent_int[0] = 0xffffffffu;
}

// One thread per sorted primitive: walks up; the second child to arrive at a node merges the two boxes.
__global__ void k_lbvh_refit(const Box6* __restrict__ boxes, const uint32_t* __restrict__ vals, int n, const uint32_t* __restrict__ left,
                             const uint32_t* __restrict__ right, const uint32_t* __restrict__ parent_int, const uint32_t* __restrict__ parent_leaf,
                             Box6* __restrict__ leaf_box, Box6* __restrict__ node_box, uint32_t* __restrict__ ticket) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    leaf_box[p] = boxes[vals[p]];
    __threadfence();
    uint32_t cur = parent_leaf[p];
    while (cur != 0xffffffffu) {
        if (atomicAdd(&ticket[cur], 1u) == 0u) return; // the sibling subtree is not finished yet
        __threadfence();
        const uint32_t lc = left[cur], rc = right[cur];
        const volatile Box6* a = (lc & LBVH_LEAF) ? &leaf_box[lc & ~LBVH_LEAF] : &node_box[lc];
        const volatile Box6* b = (rc & LBVH_LEAF) ? &leaf_box[rc & ~LBVH_LEAF] : &node_box[rc];
        Box6 u;
        for (int k = 0; k < 3; ++k) { u.lo[k] = fminf(a->lo[k], b->lo[k]); u.hi[k] = fmaxf(a->hi[k], b->hi[k]); }
        node_box[cur] = u;
        __threadfence();
        cur = parent_int[cur];
    }
}

__device__ __forceinline__ bool lbvh_collapsible(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, uint32_t i,
                                                 uint32_t max_leaf) {
    const uint32_t f = first[i], l = last[i];
    return (l - f + 1u) <= max_leaf && (keys[f] >> 60) == (keys[l] >> 60);
}

// keep[i] = 1 when internal node i stays an internal node of the emitted tree; depth of the deepest kept node.
__global__ void k_lbvh_mark(const uint64_t* __restrict__ keys, int n, const uint32_t* __restrict__ first, const uint32_t* __restrict__ last,
                            const uint32_t* __restrict__ parent_int, uint32_t max_leaf, uint32_t* __restrict__ keep, uint32_t* __restrict__ max_depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const bool k = !lbvh_collapsible(keys, first, last, (uint32_t)i, max_leaf);
    keep[i] = k ? 1u : 0u;
    if (k) {
        uint32_t depth = 1, cur = parent_int[i];
        while (cur != 0xffffffffu) { ++depth; cur = parent_int[cur]; }
        atomicMax(max_depth, depth + 1u); // + the leaf level below
    }
}

struct LbvhTypeDelta {
    long long d[8]; // typed index = sorted position + d[type]
};

__device__ __forceinline__ void lbvh_write(BvhNode32* out, const Box6& b, uint32_t first, uint32_t count) {
    BvhNode32 nd;
    for (int k = 0; k < 3; ++k) { nd.min[k] = b.lo[k]; nd.max[k] = b.hi[k]; }
    nd.first = first; nd.count = count;
    *out = nd;
}

// One thread per kept internal node: its two children become the sibling pair at node index base + 2 * (1 + rank).
__global__ void k_lbvh_emit(const uint64_t* __restrict__ keys, int n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                            const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, const uint32_t* __restrict__ keep,
                            const uint32_t* __restrict__ rank, const Box6* __restrict__ leaf_box, const Box6* __restrict__ node_box, uint32_t max_leaf,
                            uint32_t base, LbvhTypeDelta td, BvhNode32* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !keep[i]) return;
    BvhNode32* pair = out + 2u * (1u + rank[i]);
    for (int c = 0; c < 2; ++c) {
        const uint32_t ch = c ? right[i] : left[i];
        if (ch & LBVH_LEAF) {
            const uint32_t p = ch & ~LBVH_LEAF;
            const uint32_t ty = (uint32_t)(keys[p] >> 60);
            lbvh_write(pair + c, leaf_box[p], (uint32_t)((long long)p + td.d[ty]), RT_LEAF_FLAG | (ty << 24) | 1u);
        } else if (!keep[ch]) {
            const uint32_t f = first[ch], cnt = last[ch] - f + 1u;
            const uint32_t ty = (uint32_t)(keys[f] >> 60);
            lbvh_write(pair + c, node_box[ch], (uint32_t)((long long)f + td.d[ty]), RT_LEAF_FLAG | (ty << 24) | cnt);
        } else {
            lbvh_write(pair + c, node_box[ch], base + 2u * (1u + rank[ch]), 0u);
        }
    }
    if (i == 0) { // root pair: (root box -> the pair of node 0, empty leaf)
        lbvh_write(out, node_box[0], base + 2u, 0u);
        Box6 e;
        for (int k = 0; k < 3; ++k) { e.lo[k] = 3.0e38f; e.hi[k] = -3.0e38f; }
        lbvh_write(out + 1, e, 0u, RT_LEAF_FLAG);
    }
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T* as() { return static_cast<T*>(p); }
};

struct Carve { // offsets into one device allocation, 256-byte aligned
    size_t used = 0;
    size_t take(size_t bytes) { const size_t o = used; used = (used + bytes + 255) & ~(size_t)255; return o; }
};
struct View {
    char* p;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

#define LB_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

} // namespace

// Builds the tree of one instance on the device.  `h_boxes` = n outward-rounded, padded f32 boxes (lo[3], hi[3]),
// `h_types` = PrimType per primitive, `type_cursor[t]` = next free index of type t's device buffer (advanced on
// return), `base` = node index the emitted nodes will start at (even).  Returns the nodes (root pair first) and the
// primitive order of the leaves; *ok = false when the tree does not qualify (root collapsible, too deep) and the
// caller must use the host builder.
cudaError_t lbvh_build_device(const float* h_boxes, const uint8_t* h_types, uint32_t n, uint32_t max_leaf, uint32_t base, uint32_t type_cursor[PRIM_TYPE_COUNT],
                              std::vector<BvhNode32>& out_nodes, std::vector<uint32_t>& leaf_order, int* max_depth, float* ms_device, bool* ok) {
    *ok = false;
    if (n < 2u * max_leaf + 2u) return cudaSuccess;
    // centroid bounds and per-type counts on the host: one pass over data the host already has in cache
    float cmin[3] = {3.0e38f, 3.0e38f, 3.0e38f}, cmax[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    uint32_t per_type[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) {
            const float c = 0.5f * (h_boxes[6 * (size_t)i + k] + h_boxes[6 * (size_t)i + 3 + k]);
            cmin[k] = c < cmin[k] ? c : cmin[k]; cmax[k] = c > cmax[k] ? c : cmax[k];
        }
        ++per_type[h_types[i] & 7u];
    }
    LbvhTypeDelta td;
    {
        long long pos = 0;
        for (int t = 0; t < 8; ++t) {
            td.d[t] = (t < (int)PRIM_TYPE_COUNT ? (long long)type_cursor[t] : 0) - pos;
            pos += per_type[t];
        }
    }
    const float3 cm = make_float3(cmin[0], cmin[1], cmin[2]);
    const float3 ci = make_float3(cmax[0] > cmin[0] ? 1.f / (cmax[0] - cmin[0]) : 0.f, cmax[1] > cmin[1] ? 1.f / (cmax[1] - cmin[1]) : 0.f,
                                  cmax[2] > cmin[2] ? 1.f / (cmax[2] - cmin[2]) : 0.f);

    // one device allocation for every intermediate and the worst-case output (2 * n nodes): cudaMalloc dominates small builds
    const size_t ni = n - 1;
    size_t tmp_sort = 0, tmp_scan = 0;
    {
        cub::DoubleBuffer<uint64_t> kb0(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> vb0(nullptr, nullptr);
        LB_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, kb0, vb0, (int)n, 0, 63));
        LB_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)ni));
    }
    Carve cv;
    const size_t o_boxes = cv.take(sizeof(Box6) * n), o_types = cv.take(n), o_keys = cv.take(8 * (size_t)n), o_keys2 = cv.take(8 * (size_t)n),
                 o_vals = cv.take(4 * (size_t)n), o_vals2 = cv.take(4 * (size_t)n), o_left = cv.take(4 * ni), o_right = cv.take(4 * ni),
                 o_first = cv.take(4 * ni), o_last = cv.take(4 * ni), o_pint = cv.take(4 * ni), o_pleaf = cv.take(4 * (size_t)n),
                 o_lbox = cv.take(sizeof(Box6) * n), o_nbox = cv.take(sizeof(Box6) * ni), o_ticket = cv.take(4 * ni), o_keep = cv.take(4 * ni),
                 o_rank = cv.take(4 * ni), o_depth = cv.take(4), o_tmp = cv.take(tmp_sort > tmp_scan ? tmp_sort : tmp_scan),
                 o_out = cv.take(sizeof(BvhNode32) * 2 * (size_t)n);
    DevBuf arena;
    LB_CK(arena.alloc(cv.used));
    char* const A = arena.as<char>();
    const View d_boxes{A + o_boxes}, d_types{A + o_types}, d_keys{A + o_keys}, d_keys2{A + o_keys2}, d_vals{A + o_vals}, d_vals2{A + o_vals2},
        d_left{A + o_left}, d_right{A + o_right}, d_first{A + o_first}, d_last{A + o_last}, d_pint{A + o_pint}, d_pleaf{A + o_pleaf}, d_lbox{A + o_lbox},
        d_nbox{A + o_nbox}, d_ticket{A + o_ticket}, d_keep{A + o_keep}, d_rank{A + o_rank}, d_depth{A + o_depth}, d_tmp{A + o_tmp}, d_out{A + o_out};
    cub::DoubleBuffer<uint64_t> kb(d_keys.as<uint64_t>(), d_keys2.as<uint64_t>());
    cub::DoubleBuffer<uint32_t> vb(d_vals.as<uint32_t>(), d_vals2.as<uint32_t>());

    cudaEvent_t e0, e1;
    LB_CK(cudaEventCreate(&e0)); LB_CK(cudaEventCreate(&e1));
    LB_CK(cudaMemcpy(d_boxes.p, h_boxes, sizeof(Box6) * n, cudaMemcpyHostToDevice));
    LB_CK(cudaMemcpy(d_types.p, h_types, n, cudaMemcpyHostToDevice));
    LB_CK(cudaEventRecord(e0));
    const int T = 256, gn = (int)((n + T - 1) / T), gi = (int)((ni + T - 1) / T);
    k_lbvh_keys<<<gn, T>>>(d_boxes.as<Box6>(), d_types.as<uint8_t>(), n, cm, ci, kb.Current(), vb.Current());
    LB_CK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_sort, kb, vb, (int)n, 0, 63));
    const uint64_t* keys = kb.Current();
    const uint32_t* vals = vb.Current();
    k_lbvh_topology<<<gi, T>>>(keys, (int)n, d_left.as<uint32_t>(), d_right.as<uint32_t>(), d_first.as<uint32_t>(), d_last.as<uint32_t>(), d_pint.as<uint32_t>(),
                               d_pleaf.as<uint32_t>());
    LB_CK(cudaMemsetAsync(d_ticket.p, 0, 4 * ni));
    LB_CK(cudaMemsetAsync(d_depth.p, 0, 4));
    k_lbvh_refit<<<gn, T>>>(d_boxes.as<Box6>(), vals, (int)n, d_left.as<uint32_t>(), d_right.as<uint32_t>(), d_pint.as<uint32_t>(), d_pleaf.as<uint32_t>(),
                            d_lbox.as<Box6>(), d_nbox.as<Box6>(), d_ticket.as<uint32_t>());
    k_lbvh_mark<<<gi, T>>>(keys, (int)n, d_first.as<uint32_t>(), d_last.as<uint32_t>(), d_pint.as<uint32_t>(), max_leaf, d_keep.as<uint32_t>(), d_depth.as<uint32_t>());
    LB_CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_scan, d_keep.as<uint32_t>(), d_rank.as<uint32_t>(), (int)ni));
    uint32_t last_keep = 0, last_rank = 0, depth = 0;
    LB_CK(cudaMemcpy(&last_keep, d_keep.as<uint32_t>() + (ni - 1), 4, cudaMemcpyDeviceToHost));
    LB_CK(cudaMemcpy(&last_rank, d_rank.as<uint32_t>() + (ni - 1), 4, cudaMemcpyDeviceToHost));
    LB_CK(cudaMemcpy(&depth, d_depth.p, 4, cudaMemcpyDeviceToHost));
    const uint32_t kept = last_keep + last_rank;
    if (kept == 0 || depth > (uint32_t)RT_BVH_MAX_DEPTH) { // degenerate (everything in one leaf) or deeper than the traversal stack: host builder
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return cudaSuccess;
    }
    const size_t n_out = 2 * (size_t)(1 + kept);
    k_lbvh_emit<<<gi, T>>>(keys, (int)n, d_left.as<uint32_t>(), d_right.as<uint32_t>(), d_first.as<uint32_t>(), d_last.as<uint32_t>(), d_keep.as<uint32_t>(),
                           d_rank.as<uint32_t>(), d_lbox.as<Box6>(), d_nbox.as<Box6>(), max_leaf, base, td, d_out.as<BvhNode32>());
    LB_CK(cudaGetLastError());
    LB_CK(cudaEventRecord(e1));
    out_nodes.resize(n_out);
    leaf_order.resize(n);
    LB_CK(cudaMemcpy(out_nodes.data(), d_out.p, sizeof(BvhNode32) * n_out, cudaMemcpyDeviceToHost));
    LB_CK(cudaMemcpy(leaf_order.data(), vals, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    LB_CK(cudaEventElapsedTime(ms_device, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (int t = 0; t < (int)PRIM_TYPE_COUNT; ++t) type_cursor[t] += per_type[t];
    *max_depth = (int)depth;
    *ok = true;
    return cudaSuccess;
}

} // namespace rtb
