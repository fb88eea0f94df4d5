// fast_trace.cuh — the order-free closest-hit engine (static scenes, regular rays).
//
// What it computes.  BVHWrapper::hit (src/objects/bvhwrapper.rs:97-126) visits the leaves of the reference tree in DFS
// order with ONE running closest t; a primitive P is tested when its leaf-node box L(P) passes Aabb::hit
// (src/objects/bvh.rs:96-132) on (tmin, closest so far) and replaces the closest hit when its root is strictly closer.
// Write cand(P) for the root Sphere::hit / Triangle::hit select for P (sphere.rs:84-92, triangle.rs:118-121: the choice
// does not depend on the running closest t, only the acceptance does) and near(L) for the entry parameter of L's box in
// the reference's own slab arithmetic.  Inner-node tests only cull (Trav, device_math.cuh), so:
//     P is tested      <=>  L(P) is hit on (tmin, inf)  and  closest so far > max(tmin, near(L(P)))
//     P is accepted    <=>  P is tested  and  cand(P) < closest so far.
// Call P REGULAR when L(P) is hit and near(L(P)) <= cand(P).  If W is the DFS-first minimiser of cand over the regular
// primitives and no IRREGULAR primitive (cand(P) < near(L(P)): rounding put the hit in front of its own box) has
// cand(P) <= cand(W), the reference returns exactly W: everything it can have accepted before reaching W is farther,
// so W's leaf is entered and W accepted; nothing later is strictly closer (the oracle carries a CPU
// model of this search, oracle.cpp order_free_hit, which tests/test_oracle_properties.py compares with the reference-
// order traversal ray by ray).  W does not depend on the ORDER of the search, so the search runs near-first with early
// termination over a good tree of its own (fast_tree.h) instead of the reference's median split.
//
// What stays exact.  Every candidate is the reference's f64 arithmetic (the same sphere_hit_t / tri_hit_t / quad_hit_t as
// the reference-order engine); a candidate that would become the closest hit is checked against its REFERENCE leaf-node
// box in the reference's slab arithmetic: box missed => the reference never tests it => ignored; cand < near => the ray
// is handed to the reference-order kernel (retry list); otherwise it is accepted, ties going to the lower DFS rank.
// The search tree's f32 boxes and the culling rule are conservative and never decide a hit.
//
// The one assumption.  A subtree is culled when its box entry exceeds closest + margin, margin =
// 2^-20 (|o|_1 + 3 B) / |d| (B = largest scene coordinate): an irregular candidate can only be missed if its computed root
// lies more than `margin` in front of its own primitive's box.  Sphere roots err by at most 2^-25 (|oc| + r) / |d| (the
// discriminant's cancellation), quad roots by ulps of the plane equation, so the margin has a factor > 30 to spare;
// a triangle root can err by more only when |det| < ~1e-8 |e1||e2||d| (a ray within 1e-8 rad of the triangle's plane
// that still passes the u, v tests).  DESIGN.md 5.1b quantifies this; CR_RENDER_REFERENCE_ORDER / CR_TRACE_REFERENCE_ORDER
// select the reference-order engine, which needs no such assumption.
#pragma once
#include "device_math.cuh"

namespace crb {

enum : int { FS_IDLE = 0, FS_INNER = 1, FS_LEAF = 2, FS_DONE = 3, FS_RETRY = 4 };
static constexpr int FAST_STACK = 64;       // the builder bounds the tree depth (FAST_MAX_DEPTH)
static constexpr int FAST_SMEM_LEVELS = 8;  // stack levels kept in shared memory (template default); deeper ones (rare) in local memory

// Per-CTA lane table, SoA over the lanes (conflict free), as LaneSlots of the reference-order engine, plus the bottom of
// every lane's traversal stack: lanes of a warp sit at different stack depths, so a local-memory stack costs one L1
// wavefront per distinct depth and access (ncu: 52 % of the kernel's L1 data-pipe traffic); [level][lane] in shared
// memory is one conflict-free wavefront whatever the depths.
template <typename R, int BLOCK, int LEVELS = FAST_SMEM_LEVELS>
struct FastSlots {
    R best_t[BLOCK];
    R ray[6][BLOCK];  // (the f32 ray of the sphere pre-filter is converted from this one when a leaf needs it)
    // the box-test constants of the lane's ray (FastRay in 9 words: o*inv, the per-axis error band, 1/d).  They are read
    // into registers at the start of every INNER slice and are dead in the LEAF phase, so the f64 leaf arithmetic does
    // not push them into local memory (the 64-register build used to reload 4 of them per inner step)
    float fray[9][BLOCK];
    uint32_t best_ref[BLOCK], best_rank[BLOCK], my[BLOCK];
    uint32_t stk_ref[LEVELS][BLOCK];
    float stk_lo[LEVELS][BLOCK];
};

// Conservative f32 slab test of one child box: a lower bound of the entry parameter (clamped to tmin) and "certainly
// missed, or entered beyond the closest hit + margin".  Planes as single FFMAs (filter_box, device_math.cuh).  The
// rounding error of a plane distance, 1.01 |inv| (|b| + |o|) 2^-23 + 2^-22 |t| (device_math.cuh), is applied PER AXIS:
// the near planes are moved down and the far planes up by e_k = 1.3 |inv_k| (B + |o_k|) 2^-23 (folded into the FFMA's
// constant), so a direction that is almost parallel to one axis (|inv_k| huge) only loses the culling of that axis.
// (With one bound for all three axes such a ray could cull nothing and walked the whole tree.)
struct FastRay {
    float on_x, on_y, on_z;  // o*inv + e: constant of the near planes
    float of_x, of_y, of_z;  // o*inv - e: constant of the far planes
    float ax0, ax1, ay0, ay1, az0, az1;
};
__device__ __forceinline__ bool fast_child_fails(const NodeRec<float>& n, const FastRay& f, float tmin, float best_m, float& lo_lb) {
    const float nx = __fmaf_rn(n.xmin, f.ax0, __fmaf_rn(n.xmax, f.ax1, -f.on_x));
    const float fx = __fmaf_rn(n.xmax, f.ax0, __fmaf_rn(n.xmin, f.ax1, -f.of_x));
    const float ny = __fmaf_rn(n.ymin, f.ay0, __fmaf_rn(n.ymax, f.ay1, -f.on_y));
    const float fy = __fmaf_rn(n.ymax, f.ay0, __fmaf_rn(n.ymin, f.ay1, -f.of_y));
    const float nz = __fmaf_rn(n.zmin, f.az0, __fmaf_rn(n.zmax, f.az1, -f.on_z));
    const float fz = __fmaf_rn(n.zmax, f.az0, __fmaf_rn(n.zmin, f.az1, -f.of_z));
    const float lo = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float hi = fminf(fminf(fx, fy), fminf(fz, best_m));
    const float E = (fabsf(lo) + fabsf(hi)) * 4.7683716e-7f;  // the relative part of the plane error, 2^-21 (|lo| + |hi|)
    lo_lb = lo - E;                 // lower bound of the true entry parameter
    return (hi - lo) < -E;          // NaN / inf => false => the child is visited (culling must stay conservative)
}

// The reference's Aabb::hit on (tmin, +inf) for a regular ray (aabb_hit_regular, device_math.cuh): hit? and the entry
// parameter max(tmin, near) as the reference computes it.
template <typename R>
__device__ __forceinline__ bool ref_box_span(const NodeRec<R>& n, V3<R> o, V3<R> inv, R tmin, R& entry) {
    const bool px = inv.x > R(0), py = inv.y > R(0), pz = inv.z > R(0);
    const R x0 = (n.xmin - o.x) * inv.x, x1 = (n.xmax - o.x) * inv.x;
    const R y0 = (n.ymin - o.y) * inv.y, y1 = (n.ymax - o.y) * inv.y;
    const R z0 = (n.zmin - o.z) * inv.z, z1 = (n.zmax - o.z) * inv.z;
    const R lx = px ? x0 : x1, hx = px ? x1 : x0;
    const R ly = py ? y0 : y1, hy = py ? y1 : y0;
    const R lz = pz ? z0 : z1, hz = pz ? z1 : z0;
    R lo = (lx > tmin) ? lx : tmin;
    R hi = hx;
    lo = (ly > lo) ? ly : lo;
    hi = (hy < hi) ? hy : hi;
    lo = (lz > lo) ? lz : lo;
    hi = (hz < hi) ? hz : hi;
    entry = lo;
    return hi > lo;
}

// Small scenes: the whole search tree (inner records, leaf-primitive table, f32 spheres of the pre-filter) is copied into
// the CTA's shared memory once per launch (k_trace_fast_smem: ONE 896-thread CTA per SM, 176 KB of lane tables + up to
// 50 KB of tree).  A divergent node fetch then costs shared-memory latency (~30 cycles, no tag stage) instead of an L1
// hit (~40) or, for the third of the node loads that missed the L1 (ncu, book1), an L2 round trip.  Shared addresses are
// 32-bit window offsets; the records are read-only after the copy, so plain (non-volatile) ld.shared is safe.
struct SmemTree {
    uint32_t nodes, prims, spheres32;  // shared-window byte addresses
    // Byte stride of the 64 B node records.  At 64 the four LDS.128 of an INNER step address bank group (4 cur + j) mod 8:
    // for a fixed j only TWO of the eight 16 B bank groups, so the lanes of a warp that sit on different nodes conflict
    // (ncu, book1: 4.1 wavefronts per LDS.128 where the distinct addresses needed 2.6; the node fetches are 55 % of the
    // kernel's shared-memory wavefronts).  At 80 the group is (5 cur + j) mod 8, a bijection of cur mod 8: distinct nodes
    // spread over all banks, with the same base + immediate addressing (no extra instruction), for 25 % more node memory.
    uint32_t node_stride;
    // Quad records (128 B in f64) likewise: at a stride of 128 B every record starts in the same bank, so the eight LDS.128 of
    // a quad test by lanes on DIFFERENT quads conflict completely (ncu, Cornell box, profiles/trace_cornell_r02u.md: 12.6
    // wavefronts per LDS.128 where 3.2 were needed; those loads were 65 % of the kernel's 1.6 G shared-memory wavefronts, the
    // L1 data pipe at 89 %).  Stride = record + 16 B: bank group (9 quad + j) mod 8.
    uint32_t quad_stride;
    // the primitive records of the leaf tests (R precision) and, per leaf-table entry, the box of the primitive's
    // REFERENCE leaf node (the candidate confirmation): with these the kernel reads no scene data from global memory
    uint32_t spheres, tris, quads, leafbox;
};
// box of a reference leaf node, padded to whole 16 B words
template <typename R> struct __align__(16) LeafBox {
    R xmin, xmax, ymin, ymax, zmin, zmax;
    R pad[sizeof(R) == 8 ? 0 + 0 : 2];
};
template <> struct __align__(16) LeafBox<double> {
    double xmin, xmax, ymin, ymax, zmin, zmax;
};
template <typename T>
static __device__ __forceinline__ T lds_rec(uint32_t addr) {
    static_assert(sizeof(T) % 16 == 0, "record = whole 16 B words");
    T out;
    uint4* d = reinterpret_cast<uint4*>(&out);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i)
        asm("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(d[i].x), "=r"(d[i].y), "=r"(d[i].z), "=r"(d[i].w) : "r"(addr + 16u * (uint32_t)i));
    return out;
}
static __device__ __forceinline__ NodeRec<float> lds_node32(uint32_t addr) {
    NodeRec<float> n;
    asm("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=f"(n.xmin), "=f"(n.xmax), "=f"(n.ymin), "=f"(n.ymax) : "r"(addr));
    asm("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4+16];" : "=f"(n.zmin), "=f"(n.zmax), "=r"(n.left), "=r"(n.right) : "r"(addr));
    return n;
}
static __device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
static __device__ __forceinline__ SphereRec<float> lds_sphere32(uint32_t addr) {
    SphereRec<float> s;
    asm("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=f"(s.cx), "=f"(s.cy), "=f"(s.cz), "=f"(s.r) : "r"(addr));
    return s;
}

template <typename R, int BLOCK, bool SMEM, int LEVELS = FAST_SMEM_LEVELS>
struct FastTrav {
    FastSlots<R, BLOCK, LEVELS>* s;
    SmemTree tree;  // SMEM builds only
    __device__ __forceinline__ uint2 leaf_entry(const DevScene<R>& sc, uint32_t k) const {
        if constexpr (SMEM) return lds_u2(tree.prims + k * 8u);
        else return __ldg(&sc.fast_prims[k]);
    }
    __device__ __forceinline__ SphereRec<float> sphere32(const DevScene<R>& sc, uint32_t idx) const {
        if constexpr (SMEM) return lds_sphere32(tree.spheres32 + idx * 16u);
        else return ldg_rec<1>(sc.spheres32 + idx);
    }
    float best_m;    // (float) closest hit, rounded up, + margin
    float margin;
    uint32_t cur;    // inner node index, or the parked leaf word (FS_LEAF)
    int sp;
    // stack levels beyond the shared-memory part: two thread-local arrays OUTSIDE this struct (a dynamically indexed member
    // would force the whole struct, hot ray constants included, into local memory: 7 LDL per inner step, measured)
    uint32_t* deep_ref;
    float* deep_lo;
    __device__ __forceinline__ void push(uint32_t ref, float lo) {
        if (sp < LEVELS) {
            s->stk_ref[sp][threadIdx.x] = ref;
            s->stk_lo[sp][threadIdx.x] = lo;
        } else if (sp < FAST_STACK) {
            deep_ref[sp - LEVELS] = ref;
            deep_lo[sp - LEVELS] = lo;
        }
        ++sp;
    }

    __device__ __forceinline__ void init_from(const FilterRay& f, R tmax, float margin_k, float bmax) {
        const float k23 = 1.3f * 1.1920929e-7f;  // 1.3 * 2^-23
        const float ex = fabsf(f.ix) * (bmax + fabsf(f.ox)) * k23, ey = fabsf(f.iy) * (bmax + fabsf(f.oy)) * k23,
                    ez = fabsf(f.iz) * (bmax + fabsf(f.oz)) * k23;
        const int t = threadIdx.x;
        s->fray[0][t] = f.oix; s->fray[1][t] = f.oiy; s->fray[2][t] = f.oiz;
        s->fray[3][t] = ex; s->fray[4][t] = ey; s->fray[5][t] = ez;
        s->fray[6][t] = f.ix; s->fray[7][t] = f.iy; s->fray[8][t] = f.iz;
        s->best_t[t] = tmax;
        s->best_ref[t] = REF_MISS;
        s->best_rank[t] = 0xFFFFFFFFu;
        // margin = k (|o|_1 + 3 B) / |d|  (an overflowing or zero |d|^2 gives inf: no culling by distance, still correct)
        const float d2 = f.dx * f.dx + f.dy * f.dy + f.dz * f.dz;
        margin = margin_k * (fabsf(f.ox) + fabsf(f.oy) + fabsf(f.oz) + 3.0f * bmax) * rsqrtf(d2) * 1.0001f;
        if (!(margin >= 0.f)) margin = __int_as_float(0x7f800000);
        best_m = __double2float_ru((double)tmax) + margin;
        cur = 0u;
        sp = 0;
    }
    __device__ __forceinline__ void set_ray(V3<R> o, V3<R> d) {
        const int t = threadIdx.x;
        s->ray[0][t] = o.x; s->ray[1][t] = o.y; s->ray[2][t] = o.z; s->ray[3][t] = d.x; s->ray[4][t] = d.y; s->ray[5][t] = d.z;
    }
    __device__ __forceinline__ void get_ray(V3<R>& o, V3<R>& d) const {
        const int t = threadIdx.x;
        o = {s->ray[0][t], s->ray[1][t], s->ray[2][t]};
        d = {s->ray[3][t], s->ray[4][t], s->ray[5][t]};
    }
    // the lane's box-test constants, from the lane table (start of an INNER slice)
    __device__ __forceinline__ FastRay load_fray() const {
        const int t = threadIdx.x;
        FastRay fr;
        const float oix = s->fray[0][t], oiy = s->fray[1][t], oiz = s->fray[2][t];
        const float ex = s->fray[3][t], ey = s->fray[4][t], ez = s->fray[5][t];
        const float ix = s->fray[6][t], iy = s->fray[7][t], iz = s->fray[8][t];
        fr.on_x = oix + ex; fr.on_y = oiy + ey; fr.on_z = oiz + ez;
        fr.of_x = oix - ex; fr.of_y = oiy - ey; fr.of_z = oiz - ez;
        fr.ax0 = fmaxf(ix, 0.f); fr.ax1 = fminf(ix, 0.f);  // FilterRay::derive
        fr.ay0 = fmaxf(iy, 0.f); fr.ay1 = fminf(iy, 0.f);
        fr.az0 = fmaxf(iz, 0.f); fr.az1 = fminf(iz, 0.f);
        return fr;
    }
    // next node from the stack (entries whose box entry lies beyond the closest hit + margin are dropped)
    __device__ __forceinline__ int pop() {
        while (sp > 0) {
            --sp;
            const float lo = sp < LEVELS ? s->stk_lo[sp][threadIdx.x] : deep_lo[sp - LEVELS];
            if (!(lo > best_m)) {
                cur = sp < LEVELS ? s->stk_ref[sp][threadIdx.x] : deep_ref[sp - LEVELS];
                return (cur & FAST_LEAF) ? (int)FS_LEAF : (int)FS_INNER;
            }
        }
        return FS_DONE;
    }
    // INNER step: both child boxes from one 64 B record, nearer child first
    __device__ __forceinline__ int step_inner(const DevScene<R>& sc, const FastRay& fr, float tmin) {
        NodeRec<float> a, b;
        if constexpr (SMEM) {
            const uint32_t na = tree.nodes + cur * tree.node_stride;
            a = lds_node32(na);
            b = lds_node32(na + 32u);
        } else {
            const NodeRec<float>* half = reinterpret_cast<const NodeRec<float>*>(sc.fast_nodes + cur);
            a = ldg_node32(half);
            b = ldg_node32(half + 1);
        }
        float la, lb;
        const bool ha = !fast_child_fails(a, fr, tmin, best_m, la) && a.left != FAST_EMPTY;
        const bool hb = !fast_child_fails(b, fr, tmin, best_m, lb) && b.left != FAST_EMPTY;
        if (ha && hb) {
            const bool a_first = !(lb < la);
            push(a_first ? b.left : a.left, a_first ? lb : la);
            cur = a_first ? a.left : b.left;
        } else if (ha) {
            cur = a.left;
        } else if (hb) {
            cur = b.left;
        } else {
            return pop();
        }
        return (cur & FAST_LEAF) ? (int)FS_LEAF : (int)FS_INNER;
    }
    __device__ __forceinline__ uint32_t leaf_first() const { return cur & 0x0FFFFFFFu; }
    __device__ __forceinline__ uint32_t leaf_count() const { return ((cur >> 28) & 3u) + 1u; }
    // f64 path: every primitive of the parked leaf is a sphere whose f64 discriminant is certainly negative
    __device__ __forceinline__ bool leaf_certain_miss(const DevScene<R>& sc) const {
        if constexpr (sizeof(R) == 8) {
            const uint32_t first = leaf_first(), cnt = leaf_count();
            const uint32_t r0 = leaf_entry(sc, first).x, r1 = cnt > 1u ? leaf_entry(sc, first + 1u).x : r0;
            if (ref_kind(r0) != CR_PRIM_SPHERE || ref_kind(r1) != CR_PRIM_SPHERE) return false;  // before touching the lane table
            const int t = threadIdx.x;
            // FilterRay's f32 copy of the ray (make_filter_ray), rebuilt from the lane's R-precision ray
            PreRay pre;
            pre.ox = (float)s->ray[0][t]; pre.oy = (float)s->ray[1][t]; pre.oz = (float)s->ray[2][t];
            pre.dx = (float)s->ray[3][t]; pre.dy = (float)s->ray[4][t]; pre.dz = (float)s->ray[5][t];
            pre.o2 = pre.ox * pre.ox + pre.oy * pre.oy + pre.oz * pre.oz;
            if (!sphere_definite_miss(sphere32(sc, ref_index(r0)), pre)) return false;
            return cnt == 1u || sphere_definite_miss(sphere32(sc, ref_index(r1)), pre);
        } else {
            return false;
        }
    }
    // LEAF step: candidates in the reference's arithmetic; a candidate that would become the closest hit is checked
    // against its reference leaf-node box (regularity).  Returns the next state (FS_RETRY: hand the ray back).
    __device__ __forceinline__ int step_leaf(const DevScene<R>& sc, V3<R> o, V3<R> d, R tmin, R tmax) {
        const int t = threadIdx.x;
        const R a = vlen2(d);  // sphere.rs:74
        R best = s->best_t[t];
        uint32_t brank = s->best_rank[t], bref = REF_NONE;
        const uint32_t first = leaf_first(), cnt = leaf_count();
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint2 e = leaf_entry(sc, first + k);
            R c;
            if constexpr (SMEM) {
                const uint32_t kind = ref_kind(e.x), idx = ref_index(e.x);
                bool hit;
                if (kind == CR_PRIM_SPHERE) {
                    hit = sphere_hit_t(lds_rec<SphereRec<R>>(tree.spheres + idx * (uint32_t)sizeof(SphereRec<R>)), o, d, a, tmin, tmax, c);
                } else if (kind == CR_PRIM_TRIANGLE) {
                    hit = tri_hit_t(lds_rec<TriRec<R>>(tree.tris + idx * (uint32_t)sizeof(TriRec<R>)), o, d, tmin, tmax, c);
                } else {
                    R al, be;
                    hit = quad_hit_t(lds_rec<QuadRec<R>>(tree.quads + idx * tree.quad_stride), o, d, tmin, tmax, c, al, be);
                }
                if (!hit) continue;
            } else {
                if (!Trav<R, RegStore<R>, false>::test_prim(sc, e.x, o, d, a, tmin, tmax, R(0), c)) continue;
            }
            if constexpr (sizeof(R) == 8) {
                if (!(c < best || (c == best && e.y < brank))) continue;
                NodeRec<R> n;
                if constexpr (SMEM) {
                    const LeafBox<R> lb = lds_rec<LeafBox<R>>(tree.leafbox + (first + k) * (uint32_t)sizeof(LeafBox<R>));
                    n.xmin = lb.xmin; n.xmax = lb.xmax; n.ymin = lb.ymin; n.ymax = lb.ymax; n.zmin = lb.zmin; n.zmax = lb.zmax;
                } else {
                    n = ldg_rec<sizeof(NodeRec<R>) / 16>(sc.nodes + (e.y >> 1));
                }
                const V3<R> inv = {R(1) / d.x, R(1) / d.y, R(1) / d.z};  // adinv, bvh.rs:111
                R entry;
                if (!ref_box_span(n, o, inv, tmin, entry)) continue;  // the reference never tests this primitive
                if (c < entry) return FS_RETRY;                        // irregular candidate: reference order decides
            } else {
                if (!(c < best)) continue;  // f32 path: statistical parity only
            }
            best = c;
            brank = e.y;
            bref = e.x;
        }
        if (bref != REF_NONE) {
            s->best_t[t] = best;
            s->best_ref[t] = bref;
            s->best_rank[t] = brank;
            best_m = __double2float_ru((double)best) + margin;
        }
        return pop();
    }
};

// Warp-persistent driver: the same schedule as trace_persistent (device_math.cuh) — INNER slices of cheap f32 steps,
// then one LEAF phase for the parked lanes — with the exact box phase gone (the search tree never decides anything).
//   IO::count() / cursor() / filter(i,tmin,tmax) / load(i,o,d) / commit(has,i,ref,t,o,d,tm)   (commit is warp-synchronous)
// Rays the engine cannot decide (irregular rays: zero / NaN / inf components; rays that met an irregular candidate) are
// appended to retry_list; the reference-order kernel traces them afterwards.
static __device__ __forceinline__ uint32_t fast_warp_append(uint32_t* counter, bool pred) {
    const uint32_t mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

template <typename R, int BLOCK, bool SMEM, int LEVELS = FAST_SMEM_LEVELS, typename IO>
__device__ __forceinline__ void fast_trace_persistent(const DevScene<R>& sc, R tmin, R tmax, IO& io, FastSlots<R, BLOCK, LEVELS>* slots,
                                                      uint32_t* retry_list, uint32_t* retry_count, SmemTree tree = SmemTree{}) {
    const int NODE_SLICE = sc.node_slice;
    const int REFILL = sc.refill;
    const uint32_t n = io.count();
    const int lane = threadIdx.x & 31;
    const float tmin32 = (float)tmin;
    uint32_t deep_ref[FAST_STACK - LEVELS];
    float deep_lo[FAST_STACK - LEVELS];
    FastTrav<R, BLOCK, SMEM, LEVELS> tv;
    tv.s = slots;
    tv.tree = tree;
    tv.deep_ref = deep_ref;
    tv.deep_lo = deep_lo;
    tv.cur = 0u;
    tv.sp = 0;
    tv.best_m = 0.f;
    tv.margin = 0.f;
    int st = FS_IDLE;
    bool exhausted = false;
    const bool small = n <= gridDim.x * blockDim.x;
    for (;;) {
        const uint32_t walking = __ballot_sync(0xffffffffu, st == FS_INNER || st == FS_LEAF);
        const int n_free = 32 - __popc(walking);
        if ((!exhausted && n_free >= REFILL) || walking == 0u) {  // warp-uniform
            {
                V3<R> o = {R(0), R(0), R(0)}, d = {R(0), R(0), R(0)};
                uint32_t my = 0, bref = REF_MISS;
                R bt = tmax;
                if (st == FS_DONE || st == FS_RETRY) {
                    my = slots->my[threadIdx.x];
                    bref = slots->best_ref[threadIdx.x];
                    bt = slots->best_t[threadIdx.x];
                    if (io.commit_needs_ray()) tv.get_ray(o, d);
                }
                io.commit(st == FS_DONE, my, bref, bt, o, d, R(0));
                const uint32_t pos = fast_warp_append(retry_count, st == FS_RETRY);
                if (st == FS_RETRY) retry_list[pos] = io.item(my);
            }
            if (st == FS_DONE || st == FS_RETRY) st = FS_IDLE;
            if (!exhausted) {
                uint32_t base = 0;
                if (small) {
                    base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u;
                } else {
                    if (lane == 0) base = atomicAdd(io.cursor(), (uint32_t)n_free);
                    base = __shfl_sync(0xffffffffu, base, 0);
                }
                if (st == FS_IDLE) {
                    const uint32_t k = base + (uint32_t)__popc(~walking & ((1u << lane) - 1u));
                    if (k < n) {
                        slots->my[threadIdx.x] = k;
                        V3<R> o, d;
                        io.load(k, o, d);
                        const FilterRay f = io.filter(k, tmin, tmax);
                        tv.init_from(f, tmax, sc.fast_margin_k, sc.bmax);
                        tv.set_ray(o, d);
                        st = f.ok ? (int)FS_INNER : (int)FS_RETRY;  // irregular rays keep the reference's test at every node
                    }
                }
                if (small || base + (uint32_t)n_free >= n) exhausted = true;
            }
            if (__ballot_sync(0xffffffffu, st != FS_IDLE) == 0u) break;
        }
        if (__any_sync(0xffffffffu, st == FS_INNER)) {
            const FastRay fr = tv.load_fray();
#pragma unroll 1
            for (int k = 0; k < NODE_SLICE; k += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (st == FS_INNER) st = tv.step_inner(sc, fr, tmin32);
                }
                if (__popc(__ballot_sync(0xffffffffu, st == FS_INNER)) < sc.min_node_lanes) break;
            }
        }
        if (st == FS_LEAF && tv.leaf_certain_miss(sc)) st = tv.pop();
        if (__any_sync(0xffffffffu, st == FS_LEAF)) {
            if (st == FS_LEAF) {
                V3<R> o, d;
                tv.get_ray(o, d);
                st = tv.step_leaf(sc, o, d, tmin, tmax);
            }
        }
    }
}

// Closest hit for ONE ray per lane with the order-free engine, all 32 lanes of the warp together (no refill, no queues):
// the tail kernel's trace step.  Lanes with active == false only take part in the votes.  Returns the closest hit in
// (ref, t); `retry` = the lane's ray must be traced in reference order instead (irregular ray or irregular candidate).
template <typename R, int BLOCK>
__device__ __forceinline__ void fast_trace_warp_batch(const DevScene<R>& sc, R tmin, R tmax, bool active, V3<R> o, V3<R> d,
                                                      FastSlots<R, BLOCK>* slots, uint32_t& ref, R& t, bool& retry) {
    uint32_t deep_ref[FAST_STACK - FAST_SMEM_LEVELS];
    float deep_lo[FAST_STACK - FAST_SMEM_LEVELS];
    FastTrav<R, BLOCK, false> tv;
    tv.s = slots;
    tv.tree = SmemTree{};
    tv.deep_ref = deep_ref;
    tv.deep_lo = deep_lo;
    tv.cur = 0u;
    tv.sp = 0;
    tv.best_m = 0.f;
    tv.margin = 0.f;
    const float tmin32 = (float)tmin;
    int st = FS_DONE;
    if (active) {
        const V3<R> inv = {R(1) / d.x, R(1) / d.y, R(1) / d.z};  // adinv, bvh.rs:111
        const FilterRay f = make_filter_ray<R>(o, d, inv, tmin, tmax, sc.bsmall, sc.bmax);
        tv.init_from(f, tmax, sc.fast_margin_k, sc.bmax);
        tv.set_ray(o, d);
        st = f.ok ? (int)FS_INNER : (int)FS_RETRY;
    }
    while (__any_sync(0xffffffffu, st == FS_INNER || st == FS_LEAF)) {
        if (__any_sync(0xffffffffu, st == FS_INNER)) {
            const FastRay fr = tv.load_fray();
#pragma unroll 1
            for (int k = 0; k < sc.node_slice; k += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (st == FS_INNER) st = tv.step_inner(sc, fr, tmin32);
                }
                // one ray per lane and nothing to refill: keep stepping while any lane still has a cheap step
                if (!__any_sync(0xffffffffu, st == FS_INNER)) break;
            }
        }
        if (st == FS_LEAF && tv.leaf_certain_miss(sc)) st = tv.pop();
        if (st == FS_LEAF) {
            V3<R> ro, rd;
            tv.get_ray(ro, rd);
            st = tv.step_leaf(sc, ro, rd, tmin, tmax);
        }
    }
    retry = st == FS_RETRY;
    ref = REF_MISS;
    t = tmax;
    if (active && !retry) {
        ref = slots->best_ref[threadIdx.x];
        t = slots->best_t[threadIdx.x];
    }
}

}  // namespace crb
