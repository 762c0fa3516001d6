// bvh_build.hpp — host builder of the flattened device BVH.
//
// Replaces BvhNode::new (src/bvh.rs:14-83: median split on a random x/y axis over a cloned object
// vector) with a binned-SAH build over all three axes that emits the 32-byte sibling-pair layout of
// rt_types.h.  The reference's closest-hit result does not depend on tree topology
// (bvh.rs:97-112), so a different tree is a legal replacement.
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
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

// Result of one subtree built into its own buffers: node links are local (pair index into `nodes`), leaf `first`
// fields are local per-type counters; both are rebased when the subtree is spliced into the scene's node array.
struct Subtree {
    std::vector<BvhNode32> nodes;      // sibling pairs below the subtree root, local indices
    std::vector<uint32_t> leaf_order;  // prim `src` indices in leaf (depth-first) order
    uint32_t per_type[PRIM_TYPE_COUNT] = {0, 0, 0, 0, 0, 0};
    BvhNode32 root{};                  // content of the root node (lives in the parent's pair)
    int max_depth = 0;
};

class BvhBuilder {
public:
    BvhBuilder(std::vector<BvhNode32>& nodes, const BuildOptions& opt) : nodes_(nodes), opt_(opt) {}

    // `type_cursor[t]` = next free index of type t's device buffer; leaves take consecutive indices in depth-first order.
    // Large inputs: the top of the tree is built on the calling thread until the ranges are small enough, the subtrees
    // below are built by worker threads into private buffers and spliced in depth-first order, so the result does not
    // depend on the thread count or on scheduling.
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
        const uint32_t n_all = (uint32_t)prims.size();
        const bool timing = std::getenv("RTB200_COMMIT_TIMING") != nullptr && n_all >= 65536u;
        auto tp0 = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (!timing) return;
            const auto t = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[rtb200 commit]   sah %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
            tp0 = t;
        };
        unsigned hw = std::thread::hardware_concurrency();
        const unsigned threads = n_all >= 32768u ? std::min(16u, std::max(1u, hw)) : 1u;
        // ---- top of the tree (serial): split until a range is a task
        const uint32_t task_max = threads > 1 ? std::max<uint32_t>(4096u, n_all / (8u * threads)) : n_all;
        struct Task { uint32_t node, lo, hi; int depth; };
        std::vector<Task> tasks; // in depth-first order
        {
            std::vector<Task> stack;
            stack.push_back({res.root, 0u, n_all, 1});
            while (!stack.empty()) {
                const Task w = stack.back();
                stack.pop_back();
                if (w.hi - w.lo <= task_max) { tasks.push_back(w); continue; }
                BvhNode32 nd{};
                uint32_t mid = w.lo;
                const bool leaf = split(prims, w.lo, w.hi, nd, mid, threads);
                if (leaf) { tasks.push_back(w); continue; } // cannot happen above task_max >= max_leaf; handled by the task path
                const uint32_t left = (uint32_t)nodes_.size();
                nodes_.push_back(BvhNode32{});
                nodes_.push_back(BvhNode32{});
                nd.first = left;
                nd.count = 0;
                nodes_[w.node] = nd;
                res.max_depth = std::max(res.max_depth, w.depth);
                stack.push_back({left + 1, mid, w.hi, w.depth + 1}); // right first: the left subtree comes first in depth-first order
                stack.push_back({left, w.lo, mid, w.depth + 1});
            }
        }
        lap("top (serial)");
        // ---- subtrees (parallel): every task owns a disjoint range of `prims`
        std::vector<Subtree> sub(tasks.size());
        auto run = [&](size_t k) { build_subtree(prims, tasks[k].lo, tasks[k].hi, tasks[k].depth, sub[k]); };
        if (threads <= 1 || tasks.size() <= 1) {
            for (size_t k = 0; k < tasks.size(); ++k) run(k);
        } else {
            std::atomic<size_t> next{0};
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < threads; ++t)
                pool.emplace_back([&]() { for (size_t k = next.fetch_add(1); k < tasks.size(); k = next.fetch_add(1)) run(k); });
            for (std::thread& th : pool) th.join();
        }
        lap("subtrees");
        // ---- splice in depth-first order: rebase node links and typed leaf indices
        size_t total_nodes = nodes_.size(), total_leaf = 0;
        for (const Subtree& st : sub) { total_nodes += st.nodes.size(); total_leaf += st.leaf_order.size(); }
        nodes_.reserve(total_nodes);
        res.leaf_order.reserve(total_leaf);
        for (size_t k = 0; k < tasks.size(); ++k) {
            Subtree& st = sub[k];
            const uint32_t base = (uint32_t)nodes_.size();
            auto rebase = [&](BvhNode32& nd) {
                if (nd.count & RT_LEAF_FLAG) nd.first += type_cursor[(nd.count >> 24) & 0x7fu];
                else nd.first += base;
            };
            rebase(st.root);
            nodes_[tasks[k].node] = st.root;
            for (BvhNode32& nd : st.nodes) { rebase(nd); nodes_.push_back(nd); }
            for (int t = 0; t < (int)PRIM_TYPE_COUNT; ++t) type_cursor[t] += st.per_type[t];
            res.leaf_order.insert(res.leaf_order.end(), st.leaf_order.begin(), st.leaf_order.end());
            res.max_depth = std::max(res.max_depth, st.max_depth);
        }
        lap("splice");
        return res;
    }

private:
    // One node over prims[lo, hi): fills the outward-rounded box; returns true when it must be a leaf, else the split
    // position `mid` (prims partitioned in place).  Binned SAH over the three axes, all three binned in one pass.
    struct Bounds {
        double bmin[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, bmax[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        double cmin[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, cmax[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        bool homogeneous = true;
    };
    struct Bin { double mn[3], mx[3]; uint32_t n; };
    static const int NB = 16;
    struct Bins { Bin b[3][NB]; };

    static void bounds_pass(const std::vector<BuildPrim>& prims, uint32_t lo, uint32_t hi, uint32_t type0, Bounds& B) {
        for (uint32_t i = lo; i < hi; ++i) {
            const BuildPrim& p = prims[i];
            for (int a = 0; a < 3; ++a) {
                B.bmin[a] = std::min(B.bmin[a], p.bmin[a]); B.bmax[a] = std::max(B.bmax[a], p.bmax[a]);
                const double c = 0.5 * (p.bmin[a] + p.bmax[a]);
                B.cmin[a] = std::min(B.cmin[a], c); B.cmax[a] = std::max(B.cmax[a], c);
            }
            B.homogeneous = B.homogeneous && p.type == type0;
        }
    }
    static void clear_bins(Bins& Q) {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < NB; ++b) { for (int k = 0; k < 3; ++k) { Q.b[a][b].mn[k] = DBL_MAX; Q.b[a][b].mx[k] = -DBL_MAX; } Q.b[a][b].n = 0; }
    }
    static void bin_pass(const std::vector<BuildPrim>& prims, uint32_t lo, uint32_t hi, const double cmin[3], const double scale[3], const bool usable[3], Bins& Q) {
        for (uint32_t i = lo; i < hi; ++i) {
            const BuildPrim& p = prims[i];
            for (int a = 0; a < 3; ++a) {
                if (!usable[a]) continue;
                int b = (int)((0.5 * (p.bmin[a] + p.bmax[a]) - cmin[a]) * scale[a]);
                b = std::min(std::max(b, 0), NB - 1);
                Bin& q = Q.b[a][b];
                for (int k = 0; k < 3; ++k) { q.mn[k] = std::min(q.mn[k], p.bmin[k]); q.mx[k] = std::max(q.mx[k], p.bmax[k]); }
                q.n++;
            }
        }
    }
    // fn(chunk, lo, hi) over `chunks` equal slices of [lo, hi) on as many threads; min / max / count merges are order independent
    template <class F> static void chunked(uint32_t lo, uint32_t hi, unsigned chunks, F fn) {
        std::vector<std::thread> pool;
        const uint64_t n = hi - lo;
        for (unsigned c = 0; c < chunks; ++c) {
            const uint32_t a = lo + (uint32_t)(n * c / chunks), b = lo + (uint32_t)(n * (c + 1) / chunks);
            pool.emplace_back([=]() { fn(c, a, b); });
        }
        for (std::thread& th : pool) th.join();
    }

    // `par` > 1: the two read-only passes of a large range run chunked on `par` threads (top of the tree)
    bool split(std::vector<BuildPrim>& prims, uint32_t lo, uint32_t hi, BvhNode32& nd, uint32_t& mid, unsigned par = 1) const {
        Bounds B;
        const uint32_t type0 = prims[lo].type;
        if (par > 1 && hi - lo >= 65536u) {
            std::vector<Bounds> part(par);
            chunked(lo, hi, par, [&](unsigned c, uint32_t a, uint32_t b) { bounds_pass(prims, a, b, type0, part[c]); });
            for (const Bounds& q : part) {
                for (int a = 0; a < 3; ++a) {
                    B.bmin[a] = std::min(B.bmin[a], q.bmin[a]); B.bmax[a] = std::max(B.bmax[a], q.bmax[a]);
                    B.cmin[a] = std::min(B.cmin[a], q.cmin[a]); B.cmax[a] = std::max(B.cmax[a], q.cmax[a]);
                }
                B.homogeneous = B.homogeneous && q.homogeneous;
            }
        } else {
            bounds_pass(prims, lo, hi, type0, B);
        }
        const double* bmin = B.bmin; const double* bmax = B.bmax; const double* cmin = B.cmin; const double* cmax = B.cmax;
        const bool homogeneous = B.homogeneous;
        for (int a = 0; a < 3; ++a) {
            nd.min[a] = f32_floor(bmin[a] - opt_.pad);
            nd.max[a] = f32_ceil(bmax[a] + opt_.pad);
        }
        const uint32_t n = hi - lo;
        mid = lo;
        if (n == 1) return true;
        Bins Q;
        clear_bins(Q);
        double scale[3];
        bool usable[3];
        for (int a = 0; a < 3; ++a) {
            const double ext = cmax[a] - cmin[a];
            usable[a] = ext > 0.0;
            scale[a] = usable[a] ? NB / ext : 0.0;
        }
        if (par > 1 && hi - lo >= 65536u) {
            std::vector<Bins> part(par);
            chunked(lo, hi, par, [&](unsigned c, uint32_t a, uint32_t b) { clear_bins(part[c]); bin_pass(prims, a, b, cmin, scale, usable, part[c]); });
            for (const Bins& q : part)
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < NB; ++b) {
                        for (int k = 0; k < 3; ++k) { Q.b[a][b].mn[k] = std::min(Q.b[a][b].mn[k], q.b[a][b].mn[k]); Q.b[a][b].mx[k] = std::max(Q.b[a][b].mx[k], q.b[a][b].mx[k]); }
                        Q.b[a][b].n += q.b[a][b].n;
                    }
        } else {
            bin_pass(prims, lo, hi, cmin, scale, usable, Q);
        }
        Bin (&bins)[3][NB] = Q.b;
        double best_cost = DBL_MAX;
        int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            if (!usable[a]) continue;
            double right_area[NB];
            uint32_t right_n[NB];
            {
                double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
                uint32_t cnt = 0;
                for (int b = NB - 1; b >= 1; --b) {
                    for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], bins[a][b].mn[k]); mx[k] = std::max(mx[k], bins[a][b].mx[k]); }
                    cnt += bins[a][b].n;
                    right_area[b] = cnt ? area(mn, mx) : 0.0;
                    right_n[b] = cnt;
                }
            }
            double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
            uint32_t cnt = 0;
            for (int b = 0; b < NB - 1; ++b) {
                for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], bins[a][b].mn[k]); mx[k] = std::max(mx[k], bins[a][b].mx[k]); }
                cnt += bins[a][b].n;
                if (cnt == 0 || right_n[b + 1] == 0) continue;
                const double cost = area(mn, mx) * cnt + right_area[b + 1] * right_n[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
            }
        }
        const double parent_area = area(bmin, bmax);
        const double leaf_cost = opt_.cost_prim * n;
        const double split_cost = best_axis >= 0 && parent_area > 0.0 ? opt_.cost_traverse + opt_.cost_prim * best_cost / parent_area : DBL_MAX;
        if ((int)n <= opt_.max_leaf && homogeneous && leaf_cost <= split_cost) return true;
        if (best_axis >= 0) {
            const int a = best_axis;
            const double sc = scale[a], c0 = cmin[a];
            const int bb = best_bin;
            auto it = std::partition(prims.begin() + lo, prims.begin() + hi, [=](const BuildPrim& p) {
                int b = (int)((0.5 * (p.bmin[a] + p.bmax[a]) - c0) * sc);
                b = std::min(std::max(b, 0), 15);
                return b <= bb;
            });
            mid = (uint32_t)(it - prims.begin());
            if (mid == lo || mid == hi) mid = lo + n / 2;
            return false;
        }
        // coincident centroids (e.g. the dragon room's ceiling and ceiling light): split by type first so that
        // leaves stay homogeneous, else by index
        if (!homogeneous) {
            const uint32_t t0 = prims[lo].type;
            auto it = std::partition(prims.begin() + lo, prims.begin() + hi, [=](const BuildPrim& p) { return p.type == t0; });
            mid = (uint32_t)(it - prims.begin());
            return false;
        }
        if ((int)n <= opt_.max_leaf) return true;
        mid = lo + n / 2;
        return false;
    }

    void build_subtree(std::vector<BuildPrim>& prims, uint32_t lo0, uint32_t hi0, int depth0, Subtree& out) const {
        struct Work { int64_t node; uint32_t lo, hi; int depth; }; // node = -1: the subtree root, else local index
        std::vector<Work> stack;
        stack.push_back({-1, lo0, hi0, depth0});
        while (!stack.empty()) {
            const Work w = stack.back();
            stack.pop_back();
            out.max_depth = std::max(out.max_depth, w.depth);
            BvhNode32 nd{};
            uint32_t mid = w.lo;
            if (split(prims, w.lo, w.hi, nd, mid)) {
                const uint32_t type = prims[w.lo].type, n = w.hi - w.lo;
                nd.first = out.per_type[type];
                nd.count = RT_LEAF_FLAG | (type << 24) | n;
                out.per_type[type] += n;
                for (uint32_t i = w.lo; i < w.hi; ++i) out.leaf_order.push_back(prims[i].src);
            } else {
                const uint32_t left = (uint32_t)out.nodes.size();
                out.nodes.push_back(BvhNode32{});
                out.nodes.push_back(BvhNode32{});
                nd.first = left;
                nd.count = 0;
                stack.push_back({(int64_t)left + 1, mid, w.hi, w.depth + 1});
                stack.push_back({(int64_t)left, w.lo, mid, w.depth + 1});
            }
            if (w.node < 0) out.root = nd; else out.nodes[(size_t)w.node] = nd;
        }
    }

    static double area(const double mn[3], const double mx[3]) {
        const double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
    std::vector<BvhNode32>& nodes_;
    BuildOptions opt_;
};

} // namespace rtb
