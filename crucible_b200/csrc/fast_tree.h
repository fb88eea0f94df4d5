// fast_tree.h — host build of the SEARCH tree of the order-free trace engine (fast_trace.cuh).
//
// The reference's closest hit (BVHWrapper::hit, src/objects/bvhwrapper.rs:97-126) has an order-free description
// (DESIGN.md 5.1b): the winner is the DFS-first minimiser of the candidate roots over the primitives whose REFERENCE
// leaf-node box is hit, provided no "irregular" candidate interferes.  The reference tree is therefore only needed for a
// candidate's leaf-node box and DFS rank; the SEARCH may use any conservative structure over the same primitives.
// This one is a binned-SAH binary BVH over the construction-time primitive boxes (the boxes the reference builds its
// own tree from: Sphere::new sphere.rs:29-30, Triangle::new triangle.rs:27-35), leaves of 1-2 primitives, f32 boxes
// rounded outward.  On the BASELINE meshes it needs 5-6x fewer box tests per ray than the reference's median split
// (measured with the oracle's model of this search, oracle.cpp order_free_hit): it is NOT the reference's tree and
// never decides a hit by itself.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <limits>
#include <thread>
#include <vector>

#include "bvh_host.h"

namespace crb {

static constexpr int FAST_MAX_DEPTH = 62;  // the traversal stack holds 64 entries

struct FastTreeHost {
    std::vector<FastNodeRec> nodes;   // nodes[0] = root (always an inner record)
    std::vector<uint32_t> leaf_prims; // leaf-primitive table: index into the `visible` element list, leaf by leaf
    uint32_t depth = 0;
};

namespace fast_detail {

inline float f_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
inline float f_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}
inline double half_area(const Box& b) {
    const double x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
    return x * y + y * z + z * x;
}
inline void set_half(FastHalf& h, const Box& b, uint32_t child) {
    h.xmin = f_down(b.lo[0]); h.xmax = f_up(b.hi[0]);
    h.ymin = f_down(b.lo[1]); h.ymax = f_up(b.hi[1]);
    h.zmin = f_down(b.lo[2]); h.zmax = f_up(b.hi[2]);
    h.child = child;
    h.pad = 0;
}
inline void set_empty(FastHalf& h) {
    h.xmin = h.ymin = h.zmin = 3.0e38f;
    h.xmax = h.ymax = h.zmax = -3.0e38f;
    h.child = FAST_EMPTY;
    h.pad = 0;
}

struct Builder {
    const ElementVec& elements;
    const std::vector<uint32_t>& visible;
    std::vector<uint32_t> idx;      // permutation of [0, visible.size())
    std::vector<float> cx, cy, cz;  // centroids (f32 is plenty for binning)
    FastTreeHost& out;
    std::atomic<uint32_t> next_node{1};
    std::atomic<uint32_t> max_depth{0};

    const Box& box_of(uint32_t v) const { return elements[visible[v]].box; }

    // Builds the subtree of idx[lo, hi) (hi - lo >= 3 here) into node record `me`.
    void build(uint32_t me, uint32_t lo, uint32_t hi, uint32_t depth, int par) {
        uint32_t d = max_depth.load();
        while (depth > d && !max_depth.compare_exchange_weak(d, depth)) {}
        // split
        float cmin[3] = {3e38f, 3e38f, 3e38f}, cmax[3] = {-3e38f, -3e38f, -3e38f};
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t v = idx[i];
            cmin[0] = std::min(cmin[0], cx[v]); cmax[0] = std::max(cmax[0], cx[v]);
            cmin[1] = std::min(cmin[1], cy[v]); cmax[1] = std::max(cmax[1], cy[v]);
            cmin[2] = std::min(cmin[2], cz[v]); cmax[2] = std::max(cmax[2], cz[v]);
        }
        const float ex[3] = {cmax[0] - cmin[0], cmax[1] - cmin[1], cmax[2] - cmin[2]};
        const int ax = ex[0] >= ex[1] ? (ex[0] >= ex[2] ? 0 : 2) : (ex[1] >= ex[2] ? 1 : 2);
        const std::vector<float>& key = ax == 0 ? cx : (ax == 1 ? cy : cz);
        uint32_t mid = lo + (hi - lo) / 2;
        bool split_done = false;
        if (ex[ax] > 0.f && depth < (uint32_t)FAST_MAX_DEPTH - 24) {
            constexpr int NB = 16;
            Box bb[NB];
            uint32_t bc[NB] = {0};
            const float scale = (float)NB / ex[ax];
            auto bin = [&](uint32_t v) {
                int b = (int)((key[v] - cmin[ax]) * scale);
                return b < 0 ? 0 : (b >= NB ? NB - 1 : b);
            };
            for (uint32_t i = lo; i < hi; ++i) {
                const int b = bin(idx[i]);
                bb[b] = box_union(bb[b], box_of(idx[i]));
                bc[b]++;
            }
            Box la[NB];
            uint32_t lc[NB];
            Box acc;
            uint32_t cnt = 0;
            for (int b = 0; b < NB; ++b) {
                acc = box_union(acc, bb[b]);
                cnt += bc[b];
                la[b] = acc;
                lc[b] = cnt;
            }
            acc = Box();
            cnt = 0;
            double best = 1e300;
            int bs = -1;
            for (int b = NB - 1; b > 0; --b) {
                acc = box_union(acc, bb[b]);
                cnt += bc[b];
                if (lc[b - 1] == 0 || cnt == 0) continue;
                const double c = half_area(la[b - 1]) * lc[b - 1] + half_area(acc) * cnt;
                if (c < best) {
                    best = c;
                    bs = b;
                }
            }
            if (bs > 0) {
                const uint32_t m = (uint32_t)(std::partition(idx.begin() + lo, idx.begin() + hi, [&](uint32_t v) { return bin(v) < bs; }) - idx.begin());
                if (m > lo && m < hi) {
                    mid = m;
                    split_done = true;
                }
            }
        }
        if (!split_done)  // degenerate centroids, or a branch getting too deep: median split keeps the depth bounded
            std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        const uint32_t range[2][2] = {{lo, mid}, {mid, hi}};
        uint32_t child_node[2] = {0, 0};
        for (int s = 0; s < 2; ++s) {
            const uint32_t a = range[s][0], b = range[s][1];
            Box bx;
            for (uint32_t i = a; i < b; ++i) bx = box_union(bx, box_of(idx[i]));
            if (b - a <= 2) {
                set_half(out.nodes[me].c[s], bx, FAST_LEAF | ((b - a - 1) << 28) | a);  // leaf table position = position in idx
            } else {
                child_node[s] = next_node.fetch_add(1);
                set_half(out.nodes[me].c[s], bx, child_node[s]);
            }
        }
        const bool l_inner = child_node[0] != 0, r_inner = child_node[1] != 0;
        if (par > 0 && l_inner && r_inner && hi - lo > 65536) {
            std::thread th([&] { build(child_node[0], lo, mid, depth + 1, par - 1); });
            build(child_node[1], mid, hi, depth + 1, par - 1);
            th.join();
        } else {
            if (l_inner) build(child_node[0], lo, mid, depth + 1, 0);
            if (r_inner) build(child_node[1], mid, hi, depth + 1, 0);
        }
    }
};

}  // namespace fast_detail

// elements[visible[i]] are the primitives of the committed world (hidden ones already dropped).
inline void build_fast_tree(const ElementVec& elements, const std::vector<uint32_t>& visible, FastTreeHost& out) {
    using namespace fast_detail;
    const uint32_t n = (uint32_t)visible.size();
    out.nodes.clear();
    out.leaf_prims.clear();
    out.depth = 0;
    if (n == 0) return;
    Builder b{elements, visible, {}, {}, {}, {}, out};
    b.idx.resize(n);
    b.cx.resize(n);
    b.cy.resize(n);
    b.cz.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        b.idx[i] = i;
        const Box& bx = elements[visible[i]].box;
        b.cx[i] = (float)(0.5 * (bx.lo[0] + bx.hi[0]));
        b.cy[i] = (float)(0.5 * (bx.lo[1] + bx.hi[1]));
        b.cz[i] = (float)(0.5 * (bx.lo[2] + bx.hi[2]));
    }
    out.nodes.resize(std::max<uint32_t>(1u, n));  // an inner record has >= 3 primitives below it: fewer than n records
    if (n <= 2) {
        Box bx;
        for (uint32_t i = 0; i < n; ++i) bx = box_union(bx, elements[visible[i]].box);
        set_half(out.nodes[0].c[0], bx, FAST_LEAF | ((n - 1) << 28) | 0u);
        set_empty(out.nodes[0].c[1]);
        out.depth = 1;
    } else {
        b.build(0, 0, n, 1, 4);
        out.depth = b.max_depth.load() + 1;
    }
    out.nodes.resize(b.next_node.load());
    out.leaf_prims.resize(n);
    for (uint32_t i = 0; i < n; ++i) out.leaf_prims[i] = visible[b.idx[i]];
}

}  // namespace crb
