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
#pragma once
#include <cuda_runtime.h> // float4, make_float4

#include <cstdint>
#include <cstring>
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
        R.root = base;
        R.max_depth = 1;
        R.ok = true;
        return R;
    }
    struct Item { uint32_t bin_node; uint32_t wide; int depth; };
    std::vector<Item> todo;
    out.resize(out0 + 8, make_float4(0.f, 0.f, 0.f, 0.f));
    todo.push_back({root, base, 1});
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        if (it.depth > R.max_depth) R.max_depth = it.depth;
        if (it.depth > RT_WIDE_MAX_DEPTH) return R;
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
        // interior children are laid out depth-first: reserve their wide nodes now (contiguous), fill them later
        for (int k = 0; k < nk; ++k) {
            const BvhNode32& c = bin[kids[k]];
            if (c.count) {
                if (!leaf_ref(c, ref[k])) return R;
                if (ref[k] == RT_WIDE_EMPTY) continue;
            } else {
                const size_t idx = out.size() / 8;
                if (idx >= (1u << 31)) return R;
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
    }
    R.root = base;
    R.ok = true;
    return R;
}

} // namespace rtb
