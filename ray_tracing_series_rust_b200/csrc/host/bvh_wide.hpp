// bvh_wide.hpp — collapses the binary sibling-pair BVH (bvh_build.hpp / lbvh.cu) into 4-wide nodes.
//
// The reference walks a binary tree of Arc<Box<dyn Hittable>> (src/bvh.rs:97-112).  Closest hit does not depend on
// the topology (bvh.rs:97-112 vs hit.rs:660-690), so the device may walk any tree over the same leaves.  On the
// 871 200-triangle mesh the binary walk is bound by the latency of its dependent node fetches (profiles/README.md);
// a 4-wide node answers two levels of the binary tree with one 128-byte fetch.
//
// Collapse: a wide node starts from the two children of a binary node and repeatedly replaces the interior child of
// largest surface area by its two children until it has four children or only leaves (Wald et al. 2008, Dammertz et
// al. 2008).  Boxes are the binary nodes' own f32 boxes (already rounded outward and padded), so conservativeness is
// unchanged.  Layout: rt_types.h (RT_WIDE_EMPTY).
//
// Large trees: the top levels are collapsed on the calling thread, the subtrees below a fixed wide depth by worker
// threads into private buffers that are spliced in task order with their references relocated, so the result does
// not depend on the thread count or on scheduling (871 200 triangles: 105 -> 54 ms on 8 cores).
#pragma once
#include <cuda_runtime.h> // float4, make_float4

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../rt_types.h"

namespace rtb {

struct WideResult {
    uint32_t root = RT_WIDE_EMPTY; // reference of the root in `out` (absolute wide-node index)
    int max_depth = 0;  // wide nodes on the longest root-to-leaf path
    bool ok = false;    // false: a leaf does not fit the reference encoding or the tree is too deep -> keep the binary walk
};

namespace wide_detail {
inline float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline double half_area(const BvhNode32& n) {
    const double dx = (double)n.max[0] - n.min[0], dy = (double)n.max[1] - n.min[1], dz = (double)n.max[2] - n.min[2];
    return dx * dy + dy * dz + dz * dx;
}
// leaf reference of a binary leaf node, RT_WIDE_EMPTY for an empty leaf; false if it cannot be encoded
inline bool leaf_ref(const BvhNode32& n, uint32_t& ref) {
    const uint32_t type = (n.count >> 24) & 0x7fu, cnt = n.count & 0xffffffu;
    if (cnt == 0) { ref = RT_WIDE_EMPTY; return true; }
    if (cnt > 8u || type > 6u || n.first >= (1u << 25)) return false;
    ref = RT_LEAF_FLAG | (type << 28) | ((cnt - 1u) << 25) | n.first;
    return true;
}

struct Deferred { uint32_t bin_node; uint32_t parent; int slot; int depth; }; // a subtree left to a worker; patches slot `slot` of wide node `parent`

// Collapses the subtree under the interior binary node `bin_root` into `out`; wide-node indices count from out's start.
// With `deferred`, interior children of nodes at wide depth >= defer_depth are not descended into but recorded (their slot
// holds RT_WIDE_EMPTY until patched).
inline bool collapse_range(const std::vector<BvhNode32>& bin, uint32_t bin_root, int depth0, int defer_depth, std::vector<float4>& out,
                           std::vector<Deferred>* deferred, int& max_depth) {
    struct Item { uint32_t bin_node; uint32_t wide; int depth; };
    std::vector<Item> todo;
    const uint32_t first_wide = (uint32_t)(out.size() / 8);
    out.resize(out.size() + 8, make_float4(0.f, 0.f, 0.f, 0.f));
    todo.push_back({bin_root, first_wide, depth0});
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        if (it.depth > max_depth) max_depth = it.depth;
        if (it.depth > RT_WIDE_MAX_DEPTH) return false;
        uint32_t kids[4];
        int nk = 2;
        kids[0] = bin[it.bin_node].first;
        kids[1] = kids[0] + 1;
        while (nk < 4) {
            int pick = -1;
            double best = -1.0;
            for (int k = 0; k < nk; ++k)
                if (!bin[kids[k]].count) {
                    const double a = half_area(bin[kids[k]]);
                    if (a > best) { best = a; pick = k; }
                }
            if (pick < 0) break;
            const uint32_t f = bin[kids[pick]].first;
            kids[pick] = f;
            kids[nk++] = f + 1;
        }
        float lo[3][4], hi[3][4];
        uint32_t ref[4];
        for (int k = 0; k < 4; ++k) {
            for (int a = 0; a < 3; ++a) { lo[a][k] = 3.0e38f; hi[a][k] = -3.0e38f; }
            ref[k] = RT_WIDE_EMPTY;
        }
        // interior children: reserve their wide nodes now (siblings contiguous), fill them later
        for (int k = 0; k < nk; ++k) {
            const BvhNode32& c = bin[kids[k]];
            if (c.count) {
                if (!leaf_ref(c, ref[k])) return false;
                if (ref[k] == RT_WIDE_EMPTY) continue;
            } else if (deferred && it.depth >= defer_depth) {
                deferred->push_back({kids[k], it.wide, k, it.depth + 1});
            } else {
                const size_t idx = out.size() / 8;
                if (idx >= (1u << 30)) return false;
                out.resize(out.size() + 8, make_float4(0.f, 0.f, 0.f, 0.f));
                ref[k] = (uint32_t)idx;
                todo.push_back({kids[k], (uint32_t)idx, it.depth + 1});
            }
            for (int a = 0; a < 3; ++a) { lo[a][k] = c.min[a]; hi[a][k] = c.max[a]; }
        }
        float4* q = &out[8 * (size_t)it.wide];
        for (int a = 0; a < 3; ++a) {
            q[2 * a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
            q[2 * a + 1] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
        }
        q[6] = make_float4(as_float(ref[0]), as_float(ref[1]), as_float(ref[2]), as_float(ref[3]));
        uint32_t kid4[4] = {RT_WIDE_EMPTY, RT_WIDE_EMPTY, RT_WIDE_EMPTY, RT_WIDE_EMPTY}; // binary node behind every slot (for build_motion_wide)
        for (int k = 0; k < nk; ++k) kid4[k] = kids[k];
        std::memcpy(&q[7], kid4, 16);
    }
    return true;
}
} // namespace wide_detail

// `root` = index of the binary root node (first node of its pair; the second is the empty dummy leaf).  The wide nodes are
// appended to `out` (8 float4 each; several instances share one array); on failure `out` is restored.
inline WideResult collapse_to_wide(const std::vector<BvhNode32>& bin, uint32_t root, std::vector<float4>& out) {
    using namespace wide_detail;
    WideResult R;
    if (root >= bin.size()) return R;
    const size_t out0 = out.size();
    struct Restore { std::vector<float4>& v; size_t n; const bool& ok; ~Restore() { if (!ok) v.resize(n); } } restore{out, out0, R.ok};
    const uint32_t base = (uint32_t)(out0 / 8);
    if (bin[root].count) { // the whole instance is one leaf: a wide node with one child
        uint32_t ref;
        if (!leaf_ref(bin[root], ref)) return R;
        out.resize(out0 + 8, make_float4(0.f, 0.f, 0.f, 0.f));
        float4* q1 = &out[out0];
        const BvhNode32& n = bin[root];
        const float big = 3.0e38f;
        q1[0] = make_float4(n.min[0], big, big, big); q1[1] = make_float4(n.max[0], -big, -big, -big);
        q1[2] = make_float4(n.min[1], big, big, big); q1[3] = make_float4(n.max[1], -big, -big, -big);
        q1[4] = make_float4(n.min[2], big, big, big); q1[5] = make_float4(n.max[2], -big, -big, -big);
        q1[6] = make_float4(as_float(ref), as_float(RT_WIDE_EMPTY), as_float(RT_WIDE_EMPTY), as_float(RT_WIDE_EMPTY));
        const uint32_t kid4[4] = {root, RT_WIDE_EMPTY, RT_WIDE_EMPTY, RT_WIDE_EMPTY};
        std::memcpy(&q1[7], kid4, 16);
        R.root = base;
        R.max_depth = 1;
        R.ok = true;
        return R;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned threads = bin.size() >= 131072u ? std::min(16u, std::max(1u, hw)) : 1u;
    std::vector<Deferred> deferred;
    // top of the tree (wide depth <= 4: at most 256 subtrees) on this thread, numbered in out's own index space
    if (!collapse_range(bin, root, 1, 4, out, threads > 1 ? &deferred : nullptr, R.max_depth)) return R;
    if (!deferred.empty()) {
        std::vector<std::vector<float4>> sub(deferred.size());
        std::vector<int> sub_depth(deferred.size(), 0);
        std::vector<char> sub_ok(deferred.size(), 0);
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t t = next.fetch_add(1); t < deferred.size(); t = next.fetch_add(1))
                sub_ok[t] = collapse_range(bin, deferred[t].bin_node, deferred[t].depth, 0, sub[t], nullptr, sub_depth[t]) ? 1 : 0;
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(work);
        work();
        for (std::thread& th : pool) th.join();
        for (size_t t = 0; t < deferred.size(); ++t) {
            if (!sub_ok[t]) return R;
            R.max_depth = std::max(R.max_depth, sub_depth[t]);
        }
        // splice in task order (offsets are fixed by the sizes, so the copies run in parallel); references inside a subtree
        // count from its own start: relocate them
        std::vector<size_t> at(deferred.size());
        size_t total = out.size();
        for (size_t t = 0; t < deferred.size(); ++t) { at[t] = total; total += sub[t].size(); }
        if (total / 8 >= (1u << 31)) return R;
        out.resize(total);
        next = 0;
        auto splice = [&]() {
            for (size_t t = next.fetch_add(1); t < deferred.size(); t = next.fetch_add(1)) {
                const uint32_t off = (uint32_t)(at[t] / 8);
                float4* dst = &out[at[t]];
                std::memcpy(dst, sub[t].data(), sub[t].size() * sizeof(float4));
                for (size_t n = 0; n < sub[t].size(); n += 8) {
                    uint32_t refs[4];
                    std::memcpy(refs, &dst[n + 6], 16);
                    for (int k = 0; k < 4; ++k)
                        if (!(refs[k] & RT_LEAF_FLAG)) refs[k] += off;
                    std::memcpy(&dst[n + 6], refs, 16);
                }
                std::memcpy(reinterpret_cast<char*>(&out[8 * (size_t)deferred[t].parent + 6]) + 4 * deferred[t].slot, &off, 4); // parents live in the top part
                std::vector<float4>().swap(sub[t]);
            }
        };
        pool.clear();
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(splice);
        splice();
        for (std::thread& th : pool) th.join();
    }
    R.root = base;
    R.ok = true;
    return R;
}

// Motion form of the wide nodes (DeviceScene::mnodes4) from the binary motion nodes (`mbin`: node i = box at the shutter's start
// mbin[2i] and end-minus-start deltas mbin[2i + 1], scene.cpp): the same f32 values the pair walk interpolates, regrouped per wide node.
inline void build_motion_wide(const std::vector<float4>& wide, const std::vector<BvhNode32>& mbin, std::vector<float4>& out) {
    const size_t n = wide.size() / 8;
    out.assign(16 * n, make_float4(0.f, 0.f, 0.f, 0.f));
    for (size_t w = 0; w < n; ++w) {
        uint32_t kid4[4], ref4[4];
        std::memcpy(kid4, &wide[8 * w + 7], 16);
        std::memcpy(ref4, &wide[8 * w + 6], 16);
        float rows[12][4];
        for (int k = 0; k < 4; ++k) {
            const bool used = kid4[k] != RT_WIDE_EMPTY && ref4[k] != RT_WIDE_EMPTY;
            for (int a = 0; a < 3; ++a) {
                rows[2 * a][k] = used ? mbin[2 * (size_t)kid4[k]].min[a] : 3.0e38f;
                rows[2 * a + 1][k] = used ? mbin[2 * (size_t)kid4[k]].max[a] : -3.0e38f;
                rows[6 + 2 * a][k] = used ? mbin[2 * (size_t)kid4[k] + 1].min[a] : 0.f;
                rows[6 + 2 * a + 1][k] = used ? mbin[2 * (size_t)kid4[k] + 1].max[a] : 0.f;
            }
        }
        float4* q = &out[16 * w];
        for (int r = 0; r < 6; ++r) {
            q[r] = make_float4(rows[r][0], rows[r][1], rows[r][2], rows[r][3]);
            q[8 + r] = make_float4(rows[6 + r][0], rows[6 + r][1], rows[6 + r][2], rows[6 + r][3]);
        }
        q[6] = wide[8 * w + 6];
    }
}

} // namespace rtb
