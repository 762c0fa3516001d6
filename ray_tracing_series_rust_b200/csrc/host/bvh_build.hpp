// bvh_build.hpp — host builder of the flattened device BVH.
//
// Replaces BvhNode::new (src/bvh.rs:14-83: median split on a random x/y axis over a cloned object
// vector) with a binned-SAH build over all three axes that emits the 32-byte sibling-pair layout of
// rt_types.h.  The reference's closest-hit result does not depend on tree topology
// (bvh.rs:97-112), so a different tree is a legal replacement.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../rt_types.h"

namespace rtb {

struct BuildPrim {
    double bmin[3], bmax[3];
    uint32_t type; // PrimType
    uint32_t src;  // caller's index
};

struct BuildOptions {
    int max_leaf = 4;
    double cost_traverse = 1.0;
    double cost_prim = 2.0;
    double pad = 0.0; // absolute outward padding of every box (f32 slab test conservativeness)
};

struct BuildResult {
    uint32_t root = 0;
    int max_depth = 0;
    std::vector<uint32_t> leaf_order; // prim `src` indices in leaf order (grouped by leaf)
};

inline float f32_floor(double v) { // largest float <= v
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -FLT_MAX);
    return f;
}
inline float f32_ceil(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, FLT_MAX);
    return f;
}

class BvhBuilder {
public:
    BvhBuilder(std::vector<BvhNode32>& nodes, const BuildOptions& opt) : nodes_(nodes), opt_(opt) {}

    // `type_cursor[t]` = next free index of type t's device buffer; leaves take consecutive indices.
    BuildResult build(std::vector<BuildPrim>& prims, uint32_t type_cursor[PRIM_TYPE_COUNT]) {
        BuildResult res;
        if (nodes_.size() & 1u) nodes_.push_back(BvhNode32{}); // keep sibling pairs 64 B aligned
        res.root = (uint32_t)nodes_.size();
        nodes_.push_back(BvhNode32{});
        {   // the root's sibling: an empty leaf with an inverted box (never hit, nothing to test)
            BvhNode32 dummy{};
            for (int a = 0; a < 3; ++a) { dummy.min[a] = 3.0e38f; dummy.max[a] = -3.0e38f; }
            dummy.first = 0;
            dummy.count = RT_LEAF_FLAG;
            nodes_.push_back(dummy);
        }
        struct Work { uint32_t node, lo, hi; int depth; };
        std::vector<Work> stack;
        stack.push_back({res.root, 0u, (uint32_t)prims.size(), 1});
        while (!stack.empty()) {
            const Work w = stack.back();
            stack.pop_back();
            res.max_depth = std::max(res.max_depth, w.depth);
            double bmin[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, bmax[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
            double cmin[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, cmax[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
            bool homogeneous = true;
            for (uint32_t i = w.lo; i < w.hi; ++i) {
                const BuildPrim& p = prims[i];
                for (int a = 0; a < 3; ++a) {
                    bmin[a] = std::min(bmin[a], p.bmin[a]); bmax[a] = std::max(bmax[a], p.bmax[a]);
                    const double c = 0.5 * (p.bmin[a] + p.bmax[a]);
                    cmin[a] = std::min(cmin[a], c); cmax[a] = std::max(cmax[a], c);
                }
                homogeneous = homogeneous && p.type == prims[w.lo].type;
            }
            BvhNode32& nd = nodes_[w.node];
            for (int a = 0; a < 3; ++a) {
                nd.min[a] = f32_floor(bmin[a] - opt_.pad);
                nd.max[a] = f32_ceil(bmax[a] + opt_.pad);
            }
            const uint32_t n = w.hi - w.lo;
            uint32_t mid = w.lo;
            bool make_leaf = false;
            if (n == 1) {
                make_leaf = true;
            } else {
                // binned SAH over the three axes
                const int NB = 16;
                double best_cost = DBL_MAX;
                int best_axis = -1, best_bin = -1;
                for (int a = 0; a < 3; ++a) {
                    const double ext = cmax[a] - cmin[a];
                    if (!(ext > 0.0)) continue;
                    struct Bin { double mn[3], mx[3]; uint32_t n; };
                    Bin bins[NB];
                    for (int b = 0; b < NB; ++b) { for (int k = 0; k < 3; ++k) { bins[b].mn[k] = DBL_MAX; bins[b].mx[k] = -DBL_MAX; } bins[b].n = 0; }
                    const double scale = NB / ext;
                    for (uint32_t i = w.lo; i < w.hi; ++i) {
                        const BuildPrim& p = prims[i];
                        int b = (int)((0.5 * (p.bmin[a] + p.bmax[a]) - cmin[a]) * scale);
                        b = std::min(std::max(b, 0), NB - 1);
                        for (int k = 0; k < 3; ++k) { bins[b].mn[k] = std::min(bins[b].mn[k], p.bmin[k]); bins[b].mx[k] = std::max(bins[b].mx[k], p.bmax[k]); }
                        bins[b].n++;
                    }
                    double right_area[NB];
                    uint32_t right_n[NB];
                    {
                        double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
                        uint32_t cnt = 0;
                        for (int b = NB - 1; b >= 1; --b) {
                            for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], bins[b].mn[k]); mx[k] = std::max(mx[k], bins[b].mx[k]); }
                            cnt += bins[b].n;
                            right_area[b] = cnt ? area(mn, mx) : 0.0;
                            right_n[b] = cnt;
                        }
                    }
                    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
                    uint32_t cnt = 0;
                    for (int b = 0; b < NB - 1; ++b) {
                        for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], bins[b].mn[k]); mx[k] = std::max(mx[k], bins[b].mx[k]); }
                        cnt += bins[b].n;
                        if (cnt == 0 || right_n[b + 1] == 0) continue;
                        const double cost = area(mn, mx) * cnt + right_area[b + 1] * right_n[b + 1];
                        if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
                    }
                }
                const double parent_area = area(bmin, bmax);
                const double leaf_cost = opt_.cost_prim * n;
                const double split_cost = best_axis >= 0 && parent_area > 0.0 ? opt_.cost_traverse + opt_.cost_prim * best_cost / parent_area : DBL_MAX;
                if ((int)n <= opt_.max_leaf && homogeneous && leaf_cost <= split_cost) {
                    make_leaf = true;
                } else if (best_axis >= 0) {
                    const int a = best_axis;
                    const double scale = 16.0 / (cmax[a] - cmin[a]);
                    const double c0 = cmin[a];
                    const int bb = best_bin;
                    auto it = std::partition(prims.begin() + w.lo, prims.begin() + w.hi, [=](const BuildPrim& p) {
                        int b = (int)((0.5 * (p.bmin[a] + p.bmax[a]) - c0) * scale);
                        b = std::min(std::max(b, 0), 15);
                        return b <= bb;
                    });
                    mid = (uint32_t)(it - prims.begin());
                    if (mid == w.lo || mid == w.hi) mid = w.lo + n / 2;
                } else {
                    // coincident centroids (e.g. the dragon room's ceiling and ceiling light): split by
                    // type first so that leaves stay homogeneous, else by index
                    if (!homogeneous) {
                        const uint32_t t0 = prims[w.lo].type;
                        auto it = std::partition(prims.begin() + w.lo, prims.begin() + w.hi, [=](const BuildPrim& p) { return p.type == t0; });
                        mid = (uint32_t)(it - prims.begin());
                    } else if ((int)n <= opt_.max_leaf) {
                        make_leaf = true;
                    } else {
                        mid = w.lo + n / 2;
                    }
                }
            }
            if (make_leaf) {
                const uint32_t type = prims[w.lo].type;
                nd.first = type_cursor[type];
                nd.count = RT_LEAF_FLAG | (type << 24) | n;
                type_cursor[type] += n;
                for (uint32_t i = w.lo; i < w.hi; ++i) res.leaf_order.push_back(prims[i].src);
                continue;
            }
            const uint32_t left = (uint32_t)nodes_.size();
            nodes_.push_back(BvhNode32{});
            nodes_.push_back(BvhNode32{});
            BvhNode32& nd2 = nodes_[w.node]; // re-fetch: push_back may have reallocated
            nd2.first = left;
            nd2.count = 0;
            // right first so that the left subtree is processed (and its leaves numbered) first
            stack.push_back({left + 1, mid, w.hi, w.depth + 1});
            stack.push_back({left, w.lo, mid, w.depth + 1});
        }
        return res;
    }

private:
    static double area(const double mn[3], const double mx[3]) {
        const double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
    std::vector<BvhNode32>& nodes_;
    BuildOptions opt_;
};

} // namespace rtb
