// common.cuh — device-side data layout shared by every kernel of the wavefront integrator.
//
// Layout in HBM (one CrScene per device, sized for 180 GB):
//   nodes   : 64 B (f64) / 32 B (f32) per BVH node, four / two 128-bit words, preorder;
//   spheres : (cx,cy,cz,r)                 32 B / 16 B
//   tris    : (a, e1=b-a, e2=c-a)          80 B / 48 B (padded to whole 128-bit words)
//   quads   : (Q,u,v,normal,w,D)           128 B / 64 B
//   paths   : one 128 B (f64) / 64 B (f32) record per in-flight path, double buffered; the shade
//             kernels read a record through the material queue and write the survivor compactly
//             into the other buffer (stream compaction between bounces).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crucible_gpu.h"

namespace crb {

// ---- child / primitive references -------------------------------------------------------------
// bit31 = primitive; bits 29..30 = CrPrimKind; bits 0..26 = index into that kind's array (134 M).
// Node indices are plain integers.  REF_NONE marks "no right primitive" (span-1 nodes hold the same
// primitive twice in the reference, bvhwrapper.rs:59-61; the second test provably returns None).
// Node words (`left`, `right` fields of NodeRec): inner node = (skip link, split axis); leaf node =
// (left primitive ref, right primitive ref or REF_NONE).
static constexpr uint32_t REF_LEAF = 0x80000000u;
static constexpr uint32_t REF_NONE = 0x7FFFFFFFu;
static constexpr uint32_t REF_MISS = 0xFFFFFFFFu;
__host__ __device__ inline uint32_t make_leaf(uint32_t kind, uint32_t idx) { return REF_LEAF | (kind << 29) | idx; }
__host__ __device__ inline bool ref_is_leaf(uint32_t r) { return (r & REF_LEAF) != 0; }
__host__ __device__ inline uint32_t ref_kind(uint32_t r) { return (r >> 29) & 3u; }
__host__ __device__ inline uint32_t ref_index(uint32_t r) { return r & 0x07FFFFFFu; }
// bit 28 of a leaf node's first word: the node has NO box test (a member of a nested HitList, hitlist.rs:52-65: the list
// scans its objects without looking at their boxes).  The walk always "passes" such a node and tests its primitives.
static constexpr uint32_t ALWAYS_PASS_BIT = 1u << 28;
static constexpr uint32_t REF_MAX_INDEX = 0x07FFFFFEu;  // bit 27 of a node's first word = BIGBOX_BIT (device_math.cuh)

static constexpr int MAX_TEX_NEST = 8;

// ---- shading queues ------------------------------------------------------------------------------
enum Queue : int { Q_MISS = 0, Q_LAMBERTIAN = 1, Q_METAL = 2, Q_DIELECTRIC = 3, Q_EMISSIVE = 4, Q_COUNT = 5 };

// ---- node / primitive records (templated on the arithmetic type) ---------------------------------
template <typename R>
struct NodeRec;
template <>
struct __align__(16) NodeRec<double> {
    double xmin, xmax, ymin, ymax, zmin, zmax;
    uint32_t left, right, pad0, pad1;
};
template <>
struct __align__(32) NodeRec<float> {
    float xmin, xmax, ymin, ymax, zmin, zmax;
    uint32_t left, right;
};
static_assert(sizeof(NodeRec<double>) == 64, "f64 node = 4 x 128 bit");
static_assert(sizeof(NodeRec<float>) == 32, "f32 node = 2 x 128 bit");

template <typename R>
struct __align__(16) SphereRec {
    R cx, cy, cz, r;
};
template <typename R>
struct TriRec;
template <>
struct __align__(16) TriRec<double> {
    double ax, ay, az, e1x, e1y, e1z, e2x, e2y, e2z, pad;
};
template <>
struct __align__(16) TriRec<float> {
    float ax, ay, az, e1x, e1y, e1z, e2x, e2y, e2z, pad0, pad1, pad2;
};
template <typename R>
struct __align__(16) QuadRec {
    R qx, qy, qz, ux, uy, uz, vx, vy, vz, nx, ny, nz, wx, wy, wz, d;
};

struct PrimMeta {  // per primitive, per kind
    int32_t material;
    int32_t prim_index;
    int32_t obj_id;
    int32_t mat_kind;  // bits 0..7 = CrMaterialKind, bit 8 = the material's texture tree reaches an image (needs u, v)
};
static constexpr int32_t MATKIND_MASK = 0xff;
static constexpr int32_t MATKIND_NEEDS_UV = 0x100;

struct DevMaterial {
    int32_t kind, tex;
    double scatter_prob, fuzz, ior;
    double albedo[3];
    double emit[3];
};
struct DevTexture {
    int32_t kind, even, odd, image;
    double color[3];
    double inv_scale;
};
struct DevImage {
    cudaTextureObject_t tex;  // uchar4, point sampled, unnormalised coordinates
    int32_t w, h;
};

// ---- search tree of the order-free trace engine (fast_tree.h builds it, fast_trace.cuh walks it) ------------
// Device record: the boxes of BOTH children (so one 64 B load serves both tests and the near-first choice).
// child word: bit 31 clear = inner node index; bit 31 set = leaf: bits 0..27 = first entry of the leaf-primitive
// table, bits 28..29 = number of entries - 1.  FAST_EMPTY = no child (a never-hit box).
struct FastHalf {
    float xmin, xmax, ymin, ymax, zmin, zmax;
    uint32_t child, pad;
};
struct alignas(64) FastNodeRec {
    FastHalf c[2];
};
static_assert(sizeof(FastNodeRec) == 64, "fast node = 64 B");
static constexpr uint32_t FAST_LEAF = 0x80000000u;
static constexpr uint32_t FAST_EMPTY = 0xFFFFFFFFu;

// ---- object animation: evaluated keyframes (CrAnimKey) + one track (key range) per animated point --------
struct AnimTrack {
    uint32_t first, count;  // keys [first, first + count) of DevScene::anim_keys; count == 0 = static point
};

// ---- TransformTimeline::combine_and_compute for an object point (timeline/mod.rs:233-263) ---------------
// p = construction position, w = construction radius (spheres) / 1.0.  Every valid translate key adds its offset to
// its axis in list order (`translate * translate_matrix`: 1*m + 0 + 0 + v*1 = m + v exactly).  Of the scale keys only
// the LAST valid one of the list counts (`.filter(valid).next_back()`, :251-257) and its matrix multiplies the
// translated point:
//   kind 3 scale_sphere diag(1,1,1,v)  (transform_builder.rs:62-80)    -> w = v
//   kind 4 scale_x      diag(v,1,1,1)  (:146-164)                      -> x = v*x
//   kind 5 scale_y      row 1 = (v,1,0,0): the reference writes v into the WRONG slot (:229-246) -> y = v*x + y
//   kind 6 scale_z      diag(1,1,v,1)  (:312-330)                      -> z = v*z
// (rows are dot products summed left to right; the 0 * finite terms add exact zeros).  With no valid scale key the
// init matrix applies: diag(1,1,1,r) for spheres, diag(1,1,1,1) for Triangle::new's timelines (triangle.rs:24-26).
// s = clamp(proportion(t), 0, 1) keeps NaN for a zero-length interval exactly like f64::clamp.
template <typename R>
__host__ __device__ __forceinline__ void anim_eval(const CrAnimKey* keys, uint32_t first, uint32_t count, R t, R p[3], R& radius) {
    int skind = -1;
    R sv = R(0);
    for (uint32_t k = first; k < first + count; ++k) {
        const R t0 = (R)keys[k].t0, t1 = (R)keys[k].t1;
        if (!((t > t1) || (t0 <= t && t <= t1))) continue;
        R s = (t - t0) / (t1 - t0);
        s = s < R(0) ? R(0) : (s > R(1) ? R(1) : s);
        const R a = (R)keys[k].a, b = (R)keys[k].b;
        const int kind = keys[k].kind;
        if (kind < 3) {
            const R off = (keys[k].interp == CR_LERP) ? a * s : a;
            p[kind] = off + p[kind];
        } else {
            skind = kind;
            sv = (keys[k].interp == CR_LERP) ? a + (b - a) * s : b;
        }
    }
    if (skind == 3) {
        radius = sv;
    } else if (skind == 4) {
        p[0] = sv * p[0];
        radius = R(1);
    } else if (skind == 5) {
        p[1] = sv * p[0] + p[1];
        radius = R(1);
    } else if (skind == 6) {
        p[2] = sv * p[2];
        radius = R(1);
    }
}

template <typename R>
struct DevScene {
    const NodeRec<R>* nodes;
    const NodeRec<float>* nodes32;  // outward-rounded f32 copy of the boxes (conservative filter of the f64 path)
    const SphereRec<float>* spheres32;
    float bmax;                     // largest |coordinate| of the root box, rounded up
    float bsmall;                   // |coordinate| bound of the nodes whose first word has BIGBOX_BIT clear
    const SphereRec<R>* spheres;
    const TriRec<R>* tris;
    const QuadRec<R>* quads;
    const PrimMeta* meta[3];
    const DevMaterial* mats;
    const DevTexture* texs;
    const DevImage* images;
    const CrAnimKey* anim_keys;      // nullptr when nothing in the scene is animated
    const AnimTrack* sphere_track;   // [n_spheres]
    const AnimTrack* tri_track;      // [n_tris][3]: vertex timelines a, b, c
    const uint32_t* tri_anim_slot;   // [n_tris]: row of tri_anim_verts for an animated triangle
    const double* tri_anim_verts;    // [n_animated_tris][9]: construction vertices a, b, c
    const FastNodeRec* fast_nodes;  // search tree of the order-free engine (nullptr: reference-order traversal only)
    const uint2* fast_prims;        // its leaf-primitive table: (primitive ref, reference leaf node << 1 | slot = DFS rank)
    float fast_margin_k;            // culling margin factor (fast_trace.cuh)
    int32_t refill;                 // lanes idle before a warp fetches new rays
    uint32_t n_nodes;  // 0 when the world is the empty HitList (bvhwrapper.rs:29-31)
    int32_t sky_kind, sky_image;
    int32_t clamp_colors;  // 1 = reference Color semantics; 0 when the scene holds an Emissive (extension)
    int32_t node_slice;    // box tests per lane between two exact/leaf phases of the trace engine
    int32_t min_node_lanes;  // a node slice ends early when fewer lanes than this still have a cheap step
    int32_t rec_bypass_l1;   // trace kernels over a scene in global memory: path / filter records are loaded past the L1
    // an inner node the f32 filter cannot decide is entered untested (Trav::step_node) if its subtree has at most
    // free_pass_nodes nodes, or if the box is wider than free_pass_k error bands in ray parameter on every axis
    uint32_t free_pass_nodes;
    float free_pass_k;
    int32_t strict_boxes;  // scene with nested elements: every box node decides (no root skip, no free pass)
};

// ---- in-flight path record --------------------------------------------------------------------
template <typename R>
struct PathRec;
// f64: two 64 B halves = two DRAM atoms, grouped by consumer.  The first half is everything the trace kernels read (origin,
// direction, ray time); they never WRITE the record: the closest hit travels in the material queue (HitEntry below), so a
// traced ray costs one 64 B read instead of the read, a read-for-ownership of the partially written sectors and their
// write-back (ncu, round 2g: 225 B read + 75 B written per ray).  The second half is everything the miss and emissive
// shaders read — throughput, framebuffer index and a COPY of the direction (sky lookup) — so a path that leaves the scene
// costs one 64 B read; the scatter shaders read and write whole records.
template <>
struct __align__(16) PathRec<double> {
    double ox, oy, oz;               // origin
    double dx, dy, dz;               // direction (NOT normalised, ray_casting.rs:102)
    double tm;                       // ray time (ray_casting.rs:84)
    uint32_t pad0, pad1;
    double tr, tg, tb;               // product of attenuations so far
    double dx2, dy2, dz2;            // == dx, dy, dz (written together with them)
    uint32_t pixel, sample;          // GLOBAL pixel index (RNG key), sample index
    uint32_t bounce, fb;             // hits so far; local framebuffer index
};
template <>
struct __align__(16) PathRec<float> {  // 64 B = one atom
    float ox, oy, oz, tm;
    float dx, dy, dz;
    uint32_t bounce;
    float tr, tg, tb;
    uint32_t fb;
    uint32_t pixel, sample, pad0, pad1;
};
static_assert(sizeof(PathRec<double>) == 128, "f64 path record = one 128 B line");
static_assert(sizeof(PathRec<float>) == 64, "f32 path record = half a line");
// Closest hit of a path that goes to a scatter shader; entry k of material queue q sits at hits[(q - Q_LAMBERTIAN) * pool + k],
// written by the trace kernel together with the queue entry (contiguous, warp-aggregated appends: full sectors, no
// read-modify-write).  The miss and emissive queues need none.
template <typename R>
struct HitEntry;
template <>
struct __align__(16) HitEntry<double> {
    double t;
    uint32_t ref, pad;
};
template <>
struct __align__(8) HitEntry<float> {
    float t;
    uint32_t ref;
};
static constexpr int HIT_QUEUES = 3;  // lambertian, metal, dielectric

// ---- wavefront control block (device resident, one per render) -----------------------------------
struct Control {
    // per ping-pong side
    uint32_t n_in[2];        // paths to trace in buffer side s
    uint32_t out_count[2];   // survivors appended to side s by the shade kernels
    uint32_t trace_next;     // warp-level work fetch cursor of the trace kernel
    uint32_t gen_base;       // raygen: first free record of the target side
    uint32_t gen_count;      // raygen: records to generate
    uint32_t retry_count;    // order-free engine: rays handed back to the reference-order kernel this iteration
    uint64_t gen_first;      // raygen: first global sample id
    uint64_t next_sample;    // samples handed out so far
    uint64_t total_samples;  // npix_local * spp
    uint64_t rays_traced;    // sum of n_in over iterations
    uint32_t queue_count[Q_COUNT];
    uint32_t queue_next[Q_COUNT];
    uint32_t pool;
    uint32_t iteration;
    uint32_t retry_next;     // work cursor of the reference-order kernel over the retry list
    uint32_t retry_total;    // rays re-traced in reference order over the whole render (CrStats-level diagnostics)
    uint32_t tail_go;        // set by k_plan: every camera sample is issued and few paths remain: k_tail finishes them now
};

// ---- camera as the kernels see it ----------------------------------------------------------------
struct DevCamera {
    CrCamera c;
    uint32_t row_block, row_rank, row_world, rows_local;
    uint64_t seed;
    uint32_t fb_scale_bits;
    uint32_t is_static;  // no camera keyframes: raygen uses the host-evaluated basis (RaygenParams)
};

// ---- small vector helpers --------------------------------------------------------------------------
template <typename R>
struct V3 {
    R x, y, z;
};

// 128-bit read-only loads (LDG.E.128.CONSTANT) of a record made of N 16-byte words
template <int NWORDS, typename T>
__device__ __forceinline__ T ldg_rec(const T* p) {
    static_assert(sizeof(T) == NWORDS * 16, "record size");
    T out;
    const int4* s = reinterpret_cast<const int4*>(p);
    int4* d = reinterpret_cast<int4*>(&out);
#pragma unroll
    for (int i = 0; i < NWORDS; ++i) d[i] = __ldg(s + i);
    return out;
}

// One 256-bit read-only load (LDG.E.ENL2.256.CONSTANT, sm_100+) of a 32-byte f32 node: a divergent warp-wide
// load costs one L1 wavefront per distinct line touched, so one 32 B load halves the L1 data-pipe work of two
// 16 B loads of the same record (ncu: l1tex__data_pipe_lsu_wavefronts was 68 % of peak in k_trace).
__device__ __forceinline__ NodeRec<float> ldg_node32(const NodeRec<float>* p) {
#ifdef CRB_NO_LDG256
    return ldg_rec<2>(p);
#else
    NodeRec<float> n;
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(n.xmin), "=f"(n.xmax), "=f"(n.ymin), "=f"(n.ymax), "=f"(n.zmin), "=f"(n.zmax), "=r"(n.left), "=r"(n.right)
        : "l"(p));
    return n;
#endif
}

// Streamed-once loads (path / filter records at a lane refill): evict-first, so the lines of the scene stay resident.
// (ld.global.L1::no_allocate was measured too: 4 % SLOWER on book1 and Cornell — the refill's second access to the same
// record line, direction after origin, then misses.)
__device__ __forceinline__ int4 ld_stream16(const void* p) { return __ldcs(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ double ld_stream8(const double* p) { return __ldcs(p); }

}  // namespace crb
