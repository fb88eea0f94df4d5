// bvh_build.cu — device build of the reference's BVH.
//
// BVHWrapper::help_generate (src/objects/bvhwrapper.rs:46-80) is a recursive median split: the box of a span is
// the fold of its members' boxes (Aabb::new_from_boxes, bvh.rs:69-75), the span is STABLE-sorted by bbox.min on
// the box's longest axis (box_compare, bvhwrapper.rs:82-94) and cut at span/2.  The host builder in api.cu runs
// that recursion as written; this file builds the SAME tree (same nodes, same boxes bit for bit, same preorder
// indices) level by level on the device:
//
//   * All spans of one level have sizes {m, m+1} (median split), so a level is a regular grid of (span, chunk)
//     warp tasks; every node's preorder index and skip link follow from node_count() of at most four sizes per
//     level, which the host passes by value.
//   * "Stable sort of every span of the level" is ONE radix sort of the whole order array with the key
//     (span start << rbits) | rank_axis(member): span starts keep the spans where they are, the rank orders the
//     members, and the sort's stability supplies the reference's tie rule (equal keys keep their current order).
//     rank_axis = dense rank of bbox.min[axis] over the whole scene (three sorts, once), with -0.0 == +0.0 as in
//     f64::partial_cmp, so the per-level key is 2*ceil(log2 n) bits instead of 64 + 32.
//   * The box fold `r = (r <= x) ? r : x` keeps the FIRST minimum in span order, which only matters for the sign
//     of a zero; the reduction carries the position to reproduce even that.
//
// The radix sort and the prefix sum are CUB's (a library primitive, like cuBLAS for a plain GEMM); everything
// else is below.  NaN coordinates have no total order: the caller keeps such scenes on the host builder.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <chrono>
#include <map>

#include "bvh_host.h"
#include "h2d_staged.h"

namespace crb {
namespace {

constexpr int CHUNK = 1024;       // members one warp folds
constexpr int WARPS_PER_CTA = 8;  // 256 threads

struct DevBox {
    double lo[3], hi[3];
};
static_assert(sizeof(DevBox) == 48, "three 128-bit words");

struct DevNode {  // FlatNode, field for field
    double lo[3], hi[3];
    uint32_t left, right, axis, lchild, rchild, skip;
};
static_assert(sizeof(DevNode) == sizeof(FlatNode), "the device writes FlatNode records");

struct Seg {
    uint32_t start, end, at;  // members order[start, end), preorder index of the node; start == end: empty slot
};

struct Partial {  // fold of one chunk: value and position of the first extremum, per box plane
    double v[6];
    uint32_t pos[6];
};

struct LevelConsts {
    uint32_t m_hi;      // largest span of the level (spans are m_hi or m_hi - 1)
    uint32_t nc_hi;     // node_count(m_hi)
    uint32_t nc_lo;     // node_count(m_hi - 1)
    uint32_t c0;        // smallest child span of the level
    uint32_t nc_c0;     // node_count(c0)
    uint32_t nc_c1;     // node_count(c0 + 1)
    uint32_t chunks;    // (span, chunk) tasks per span
    uint32_t group;     // lanes that share one task: 32, or the power of two >= m_hi for the small spans of the deep levels
    uint32_t rbits;     // bits of a rank
};

// f64 -> u64 with the order of partial_cmp; -0.0 and +0.0 compare Equal there, so both map to +0.0's image
__device__ __forceinline__ uint64_t order_key(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    if ((b << 1) == 0) b = 0;
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// lo planes keep the first minimum, hi planes the first maximum (Interval::tight_enclose, utils.rs:631-635,
// folded left to right from Interval::EMPTY)
struct Fold {
    double v[6];
    uint32_t pos[6];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = __longlong_as_double(0x7ff0000000000000ll);
            v[3 + k] = __longlong_as_double(0xfff0000000000000ll);
            pos[k] = pos[3 + k] = 0xffffffffu;
        }
    }
    __device__ void add(int k, double x, uint32_t p) {  // k < 3: min, else max
        const bool better = (k < 3) ? (x < v[k]) : (x > v[k]);
        if (better || (x == v[k] && p < pos[k])) {
            v[k] = x;
            pos[k] = p;
        }
    }
    // reduction over groups of `group` consecutive lanes (a power of two); the result is in the group's first lane
    __device__ void group_reduce(uint32_t group) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            if ((uint32_t)off < group) {  // warp-uniform
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const double ov = __shfl_down_sync(0xffffffffu, v[k], off, group);
                    const uint32_t op = __shfl_down_sync(0xffffffffu, pos[k], off, group);
                    add(k, ov, op);
                }
            }
        }
    }
};

// Aabb::longest_axis, bvh.rs:82-94
__device__ __forceinline__ int dev_longest_axis(const double* v) {
    const double sx = v[3] - v[0], sy = v[4] - v[1], sz = v[5] - v[2];
    if (sx > sy) return (sx > sz) ? 0 : 2;
    if (sy > sz) return 1;
    return 2;
}

// one node of the level: record, children slots, sort axis
__device__ void emit_node(uint32_t slot, const Seg seg, const double* v, const LevelConsts lc, const uint32_t* __restrict__ order,
                          const uint32_t* __restrict__ leafref, DevNode* __restrict__ nodes, Seg* __restrict__ next, uint8_t* __restrict__ seg_axis) {
    const uint32_t span = seg.end - seg.start;
    DevNode n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        n.lo[k] = v[k];
        n.hi[k] = v[3 + k];
    }
    const int axis = dev_longest_axis(v);
    n.axis = (uint32_t)axis;
    n.skip = seg.at + (span == lc.m_hi ? lc.nc_hi : lc.nc_lo);
    n.lchild = n.rchild = REF_NONE;
    Seg l = {0u, 0u, 0u}, r = {0u, 0u, 0u};
    uint8_t ax = 0xff;
    if (span == 1) {
        n.left = leafref[order[seg.start]];
        n.right = REF_NONE;
    } else if (span == 2) {
        n.left = leafref[order[seg.start]];
        n.right = leafref[order[seg.start + 1]];
    } else {
        const uint32_t mid = seg.start + span / 2;
        const uint32_t lspan = mid - seg.start;
        const uint32_t l_at = seg.at + 1;
        const uint32_t r_at = l_at + (lspan == lc.c0 ? lc.nc_c0 : lc.nc_c1);
        n.left = n.lchild = l_at;
        n.right = n.rchild = r_at;
        l = {seg.start, mid, l_at};
        r = {mid, seg.end, r_at};
        ax = (uint8_t)axis;
    }
    nodes[seg.at] = n;
    next[2 * slot] = l;
    next[2 * slot + 1] = r;
    seg_axis[slot] = ax;
}

// task (slot, chunk), run by `group` lanes: fold the boxes of order[chunk]; with one chunk per span the node is emitted here
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_bvh_fold(const Seg* __restrict__ segs, uint32_t n_slots, const LevelConsts lc,
                                                                  const uint32_t* __restrict__ order, const DevBox* __restrict__ boxes,
                                                                  const uint32_t* __restrict__ leafref, Partial* __restrict__ partials,
                                                                  DevNode* __restrict__ nodes, Seg* __restrict__ next, uint8_t* __restrict__ seg_axis) {
    const uint32_t group = lc.group, per_warp = 32u / group;
    const uint64_t warp_id = (uint64_t)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31, sub = lane & (group - 1u);
    const uint64_t n_tasks = (uint64_t)n_slots * lc.chunks;
    if (warp_id * per_warp >= n_tasks) return;  // the whole warp (the shuffles below need every lane of a live warp)
    const uint64_t task = warp_id * per_warp + lane / group;
    const bool live = task < n_tasks;
    const uint32_t slot = live ? (uint32_t)(task / lc.chunks) : 0u, chunk = live ? (uint32_t)(task % lc.chunks) : 0u;
    const Seg seg = live ? segs[slot] : Seg{0u, 0u, 0u};
    const bool empty = seg.start == seg.end;
    Fold f;
    f.init();
    if (!empty) {
        const uint64_t lo = (uint64_t)seg.start + (uint64_t)chunk * CHUNK;
        const uint64_t hi = min((uint64_t)seg.end, lo + CHUNK);
        for (uint64_t p = lo + sub; p < hi; p += group) {
            const uint32_t e = order[p];
            const double2* b = reinterpret_cast<const double2*>(boxes + e);
            const double2 w0 = __ldg(b), w1 = __ldg(b + 1), w2 = __ldg(b + 2);
            f.add(0, w0.x, (uint32_t)p);
            f.add(1, w0.y, (uint32_t)p);
            f.add(2, w1.x, (uint32_t)p);
            f.add(3, w1.y, (uint32_t)p);
            f.add(4, w2.x, (uint32_t)p);
            f.add(5, w2.y, (uint32_t)p);
        }
    }
    f.group_reduce(group);
    if (sub != 0u || !live) return;
    if (empty) {
        if (lc.chunks == 1) {
            next[2 * slot] = Seg{0u, 0u, 0u};
            next[2 * slot + 1] = Seg{0u, 0u, 0u};
            seg_axis[slot] = 0xff;
        }
        return;
    }
    if (lc.chunks == 1) {
        emit_node(slot, seg, f.v, lc, order, leafref, nodes, next, seg_axis);
    } else {
        Partial pr;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            pr.v[k] = f.v[k];
            pr.pos[k] = f.pos[k];
        }
        partials[task] = pr;
    }
}

// warp per slot: fold the chunk partials (several chunks per span), emit the node
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_bvh_emit(const Seg* __restrict__ segs, uint32_t n_slots, const LevelConsts lc,
                                                                  const uint32_t* __restrict__ order, const uint32_t* __restrict__ leafref,
                                                                  const Partial* __restrict__ partials, DevNode* __restrict__ nodes,
                                                                  Seg* __restrict__ next, uint8_t* __restrict__ seg_axis) {
    const uint32_t slot = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (slot >= n_slots) return;
    const Seg seg = segs[slot];
    if (seg.start == seg.end) {
        if (lane == 0) {
            next[2 * slot] = Seg{0u, 0u, 0u};
            next[2 * slot + 1] = Seg{0u, 0u, 0u};
            seg_axis[slot] = 0xff;
        }
        return;
    }
    Fold f;
    f.init();
    const uint32_t used = (seg.end - seg.start + CHUNK - 1) / CHUNK;  // chunks of this span that hold members
    for (uint32_t c = lane; c < used; c += 32) {
        const Partial pr = partials[(uint64_t)slot * lc.chunks + c];
#pragma unroll
        for (int k = 0; k < 6; ++k) f.add(k, pr.v[k], pr.pos[k]);
    }
    f.group_reduce(32u);
    if (lane == 0) emit_node(slot, seg, f.v, lc, order, leafref, nodes, next, seg_axis);
}

// keys of the level's sort: positions outside a splitting span stay where they are
__global__ void k_bvh_identity_keys(uint64_t* __restrict__ keys, uint32_t n, uint32_t rbits) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) keys[p] = (uint64_t)p << rbits;
}
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_bvh_span_keys(const Seg* __restrict__ segs, const uint8_t* __restrict__ seg_axis, uint32_t n_slots,
                                                                       const LevelConsts lc, const uint32_t* __restrict__ order,
                                                                       const uint32_t* __restrict__ rank, uint32_t n, uint64_t* __restrict__ keys) {
    const uint32_t group = lc.group, per_warp = 32u / group;
    const uint64_t warp_id = (uint64_t)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31, sub = lane & (group - 1u);
    const uint64_t task = warp_id * per_warp + lane / group;
    if (task >= (uint64_t)n_slots * lc.chunks) return;
    const uint32_t slot = (uint32_t)(task / lc.chunks), chunk = (uint32_t)(task % lc.chunks);
    const uint32_t axis = seg_axis[slot];
    if (axis > 2) return;  // empty slot or leaf node
    const Seg seg = segs[slot];
    const uint32_t* __restrict__ rk = rank + (uint64_t)axis * n;
    const uint64_t lo = (uint64_t)seg.start + (uint64_t)chunk * CHUNK;
    const uint64_t hi = min((uint64_t)seg.end, lo + CHUNK);
    const uint64_t base = (uint64_t)seg.start << lc.rbits;
    for (uint64_t p = lo + sub; p < hi; p += group) keys[p] = base | rk[order[p]];
}

// ---- dense ranks of bbox.min per axis (once per build) ----
__global__ void k_bvh_axis_keys(const DevBox* __restrict__ boxes, uint32_t n, int axis, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = order_key(boxes[i].lo[axis]);
        vals[i] = i;
    }
}
__global__ void k_bvh_heads(const uint64_t* __restrict__ sorted, uint32_t n, uint32_t* __restrict__ head) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) head[j] = (j > 0 && sorted[j] != sorted[j - 1]) ? 1u : 0u;
}
__global__ void k_bvh_scatter_rank(const uint32_t* __restrict__ scan, const uint32_t* __restrict__ vals, uint32_t n, uint32_t* __restrict__ rank) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) rank[vals[j]] = scan[j];
}
__global__ void k_bvh_iota(uint32_t* __restrict__ a, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}


struct DevElement {  // Element (bvh_host.h), field for field: the staging array goes to the device as it is
    uint32_t kind, idx;
    uint8_t hide, pad[7];
    double lo[3], hi[3];
};
static_assert(sizeof(DevElement) == sizeof(Element) && sizeof(Element) == 64, "Element is uploaded raw");

// boxes and leaf references in visible order (`visible` == nullptr: nothing is hidden)
__global__ void k_bvh_gather(const DevElement* __restrict__ el, const uint32_t* __restrict__ visible, uint32_t n, DevBox* __restrict__ boxes,
                             uint32_t* __restrict__ leafref) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevElement e = el[visible ? visible[i] : i];
    DevBox b;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b.lo[k] = e.lo[k];
        b.hi[k] = e.hi[k];
    }
    boxes[i] = b;
    leafref[i] = make_leaf(e.kind, e.idx);
}

// BVHWrapper::new_from_vec (bvhwrapper.rs:34-44): the root box is re-derived from its two children
// (Aabb::new_from_boxes on two boxes: `<=` / `>=` keep the LEFT operand on a tie)
__global__ void k_bvh_root_box(DevNode* __restrict__ nodes, const uint32_t* __restrict__ order, const DevBox* __restrict__ boxes) {
    DevNode r = nodes[0];
    double l[6], q[6];
    if (ref_is_leaf(r.left)) {
        const DevBox a = boxes[order[0]];
        const DevBox b = (r.right == REF_NONE) ? a : boxes[order[1]];
        for (int k = 0; k < 3; ++k) {
            l[k] = a.lo[k]; l[3 + k] = a.hi[k];
            q[k] = b.lo[k]; q[3 + k] = b.hi[k];
        }
    } else {
        const DevNode a = nodes[r.left], b = nodes[r.right];
        for (int k = 0; k < 3; ++k) {
            l[k] = a.lo[k]; l[3 + k] = a.hi[k];
            q[k] = b.lo[k]; q[3 + k] = b.hi[k];
        }
    }
    for (int k = 0; k < 3; ++k) {
        nodes[0].lo[k] = (l[k] <= q[k]) ? l[k] : q[k];
        nodes[0].hi[k] = (l[3 + k] >= q[3 + k]) ? l[3 + k] : q[3 + k];
    }
}

// ---- flatten: FlatNode -> the trace kernels' records (what upload_scene does on the host for host-built trees) ----
// per-node |coordinate| bound, rounded up to f32
__global__ void k_node_bound(const DevNode* __restrict__ nodes, uint32_t n, float* __restrict__ nb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double bm = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) bm = fmax(bm, fmax(fabs(nodes[i].lo[k]), fabs(nodes[i].hi[k])));
    nb[i] = __double2float_ru(bm);
}
__global__ void k_flatten_nodes(const DevNode* __restrict__ nodes, const float* __restrict__ nb, uint32_t n, float bsmall,
                                NodeRec<double>* __restrict__ n64, NodeRec<float>* __restrict__ n32) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevNode f = nodes[i];
    // device words: inner = (skip link, axis), leaf node = (left primitive, right primitive); bit 27 = BIGBOX_BIT
    const bool leafnode = ref_is_leaf(f.left);
    const uint32_t big = nb[i] > bsmall ? (1u << 27) : 0u;
    NodeRec<double> a;
    a.xmin = f.lo[0]; a.xmax = f.hi[0];
    a.ymin = f.lo[1]; a.ymax = f.hi[1];
    a.zmin = f.lo[2]; a.zmax = f.hi[2];
    a.left = (leafnode ? f.left : f.skip) | big;
    a.right = leafnode ? f.right : f.axis;
    a.pad0 = a.pad1 = 0;
    NodeRec<float> b;  // boxes rounded OUTWARD
    b.xmin = __double2float_rd(f.lo[0]); b.xmax = __double2float_ru(f.hi[0]);
    b.ymin = __double2float_rd(f.lo[1]); b.ymax = __double2float_ru(f.hi[1]);
    b.zmin = __double2float_rd(f.lo[2]); b.zmax = __double2float_ru(f.hi[2]);
    b.left = a.left;
    b.right = a.right;
    n64[i] = a;
    n32[i] = b;
}
// triangles: a, e1 = b - a, e2 = c - a (triangle.rs:99-100); the f32 record is the rounded f64 one
__global__ void k_flatten_tris(const double* __restrict__ abc, uint32_t n, TriRec<double>* __restrict__ t64, TriRec<float>* __restrict__ t32) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* p = abc + (size_t)9 * i;
    TriRec<double> t;
    t.ax = p[0]; t.ay = p[1]; t.az = p[2];
    t.e1x = p[3] - p[0]; t.e1y = p[4] - p[1]; t.e1z = p[5] - p[2];
    t.e2x = p[6] - p[0]; t.e2y = p[7] - p[1]; t.e2z = p[8] - p[2];
    t.pad = 0.0;
    TriRec<float> u;
    u.ax = (float)t.ax; u.ay = (float)t.ay; u.az = (float)t.az;
    u.e1x = (float)t.e1x; u.e1y = (float)t.e1y; u.e1z = (float)t.e1z;
    u.e2x = (float)t.e2x; u.e2y = (float)t.e2y; u.e2z = (float)t.e2z;
    u.pad0 = u.pad1 = u.pad2 = 0.f;
    t64[i] = t;
    t32[i] = u;
}

struct NodeCounter {  // node_count with a memo: a level needs at most four sizes, each the half of an earlier one
    std::map<uint64_t, uint64_t> memo;
    uint64_t operator()(uint64_t span) {
        if (span <= 2) return 1;
        auto it = memo.find(span);
        if (it != memo.end()) return it->second;
        const uint64_t v = 1 + (*this)(span / 2) + (*this)(span - span / 2);
        memo[span] = v;
        return v;
    }
};

struct DeviceBuffers {
    cudaStream_t stream;
    std::vector<void*> ptrs;
    ~DeviceBuffers() {
        for (void* p : ptrs) cudaFreeAsync(p, stream);
    }
    template <typename T>
    cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), stream);
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return e;
    }
};

uint32_t ceil_log2(uint64_t n) {
    uint32_t b = 0;
    while (((uint64_t)1 << b) < n) ++b;
    return b;
}

}  // namespace

#define BVH_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            err = std::string("bvh device build: " #call ": ") + cudaGetErrorString(e__); \
            return CR_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

int gpu_build_bvh(int device, void* cuda_stream, const ElementVec& elements, const std::vector<uint32_t>& visible,
                  void** d_nodes_out, uint64_t* n_nodes_out, uint32_t& max_depth, BvhBuildTimes* times, std::string& err) {
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    const uint64_t n64 = visible.size();
    *d_nodes_out = nullptr;
    *n_nodes_out = 0;
    max_depth = 0;
    if (n64 == 0) return CR_OK;
    if (n64 > REF_MAX_INDEX) {
        err = "BVH too large";
        return CR_ERR_LIMIT;
    }
    const uint32_t n = (uint32_t)n64;
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    BVH_CUDA(cudaSetDevice(device));

    NodeCounter node_count_memo;
    const uint64_t total_nodes = node_count_memo(n);
    const uint32_t rbits = ceil_log2(n), sbits = ceil_log2(n);
    // slots of the widest level: spans halve until they are <= 2
    uint32_t levels = 1;
    for (uint64_t m = n; m > 2; m = (m + 1) / 2) ++levels;
    const uint64_t max_slots = (uint64_t)1 << (levels - 1);
    const uint64_t max_tasks = 3 * (n64 / CHUNK) + 64;  // only levels with several chunks per span write partials

    DeviceBuffers buf{stream, {}};
    DevElement* d_elements;
    DevBox* d_boxes;
    uint32_t *d_visible = nullptr, *d_leaf, *d_order[2], *d_rank, *d_scan, *d_vals;
    uint64_t* d_keys[2];
    DevNode* d_nodes = nullptr;
    Seg* d_segs[2];
    uint8_t* d_axis;
    Partial* d_partials;
    void* d_temp;
    const bool all_visible = visible.size() == elements.size();
    BVH_CUDA(buf.alloc(&d_elements, elements.size()));
    if (!all_visible) BVH_CUDA(buf.alloc(&d_visible, n));
    BVH_CUDA(buf.alloc(&d_boxes, n));
    BVH_CUDA(buf.alloc(&d_leaf, n));
    BVH_CUDA(buf.alloc(&d_order[0], n));
    BVH_CUDA(buf.alloc(&d_order[1], n));
    BVH_CUDA(buf.alloc(&d_rank, (size_t)3 * n));
    BVH_CUDA(buf.alloc(&d_scan, n));
    BVH_CUDA(buf.alloc(&d_vals, n));
    BVH_CUDA(buf.alloc(&d_keys[0], n));
    BVH_CUDA(buf.alloc(&d_keys[1], n));
    BVH_CUDA(buf.alloc(&d_segs[0], 2 * max_slots));
    BVH_CUDA(buf.alloc(&d_segs[1], 2 * max_slots));
    BVH_CUDA(buf.alloc(&d_axis, max_slots));
    BVH_CUDA(buf.alloc(&d_partials, max_tasks));
    size_t temp_sort = 0, temp_scan = 0;
    BVH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_sort, d_keys[0], d_keys[1], d_order[0], d_order[1], (int64_t)n, 0, 64, stream));
    BVH_CUDA(cub::DeviceScan::InclusiveSum(nullptr, temp_scan, d_scan, d_scan, (int64_t)n, stream));
    size_t temp_bytes = std::max(temp_sort, temp_scan);
    BVH_CUDA(buf.alloc(reinterpret_cast<uint8_t**>(&d_temp), temp_bytes));
    // the result outlives this call (the caller owns it)
    BVH_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_nodes), total_nodes * sizeof(DevNode), stream));
    struct NodesGuard {
        DevNode*& p;
        cudaStream_t st;
        bool keep = false;
        ~NodesGuard() {
            if (!keep && p) cudaFreeAsync(p, st);
        }
    } guard{d_nodes, stream};
    const auto t_alloc = clk::now();

    // the staging arrays go up as they are (pageable: the copies return once the bytes are staged)
    BVH_CUDA(StagedCopier::copy(device, d_elements, elements.data(), elements.size() * sizeof(Element), stream));  // 640 MB for 10 M triangles
    if (!all_visible) BVH_CUDA(cudaMemcpyAsync(d_visible, visible.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    const uint32_t tpb = 256, grid_n = (n + tpb - 1) / tpb;
    k_bvh_gather<<<grid_n, tpb, 0, stream>>>(d_elements, d_visible, n, d_boxes, d_leaf);
    BVH_CUDA(cudaStreamSynchronize(stream));
    const auto t_h2d = clk::now();

    // ---- dense ranks per axis ----
    for (int axis = 0; axis < 3; ++axis) {
        k_bvh_axis_keys<<<grid_n, tpb, 0, stream>>>(d_boxes, n, axis, d_keys[0], d_order[0]);
        size_t tb = temp_bytes;
        BVH_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, tb, d_keys[0], d_keys[1], d_order[0], d_vals, (int64_t)n, 0, 64, stream));
        k_bvh_heads<<<grid_n, tpb, 0, stream>>>(d_keys[1], n, d_scan);
        tb = temp_bytes;
        BVH_CUDA(cub::DeviceScan::InclusiveSum(d_temp, tb, d_scan, d_scan, (int64_t)n, stream));
        k_bvh_scatter_rank<<<grid_n, tpb, 0, stream>>>(d_scan, d_vals, n, d_rank + (size_t)axis * n);
    }
    k_bvh_iota<<<grid_n, tpb, 0, stream>>>(d_order[0], n);
    const Seg root = {0u, n, 0u};
    BVH_CUDA(cudaMemcpyAsync(d_segs[0], &root, sizeof(Seg), cudaMemcpyHostToDevice, stream));
    BVH_CUDA(cudaGetLastError());

    // ---- levels ----
    int cur = 0, ord = 0;
    uint64_t n_slots = 1, m_hi = n;
    uint32_t depth = 0;
    for (;;) {
        ++depth;
        LevelConsts lc;
        lc.m_hi = (uint32_t)m_hi;
        lc.nc_hi = (uint32_t)node_count_memo(m_hi);
        lc.nc_lo = (uint32_t)node_count_memo(m_hi - 1);
        lc.c0 = (uint32_t)((m_hi - 1) / 2);
        lc.nc_c0 = (uint32_t)node_count_memo(lc.c0);
        lc.nc_c1 = (uint32_t)node_count_memo((uint64_t)lc.c0 + 1);
        lc.chunks = (uint32_t)((m_hi + CHUNK - 1) / CHUNK);
        lc.group = 32;
        while (lc.group > 2 && lc.group / 2 >= m_hi) lc.group /= 2;  // deep levels: several small spans per warp
        lc.rbits = rbits;
        const uint64_t tasks = n_slots * lc.chunks;
        const uint64_t warps = (tasks + 32 / lc.group - 1) / (32 / lc.group);
        const uint32_t grid_tasks = (uint32_t)((warps + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
        k_bvh_fold<<<grid_tasks, WARPS_PER_CTA * 32, 0, stream>>>(d_segs[cur], (uint32_t)n_slots, lc, d_order[ord], d_boxes, d_leaf, d_partials, d_nodes,
                                                                  d_segs[cur ^ 1], d_axis);
        if (lc.chunks > 1) {
            const uint32_t grid_slots = (uint32_t)((n_slots + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
            k_bvh_emit<<<grid_slots, WARPS_PER_CTA * 32, 0, stream>>>(d_segs[cur], (uint32_t)n_slots, lc, d_order[ord], d_leaf, d_partials, d_nodes,
                                                                      d_segs[cur ^ 1], d_axis);
        }
        if (m_hi <= 2) break;
        k_bvh_identity_keys<<<grid_n, tpb, 0, stream>>>(d_keys[0], n, rbits);
        k_bvh_span_keys<<<grid_tasks, WARPS_PER_CTA * 32, 0, stream>>>(d_segs[cur], d_axis, (uint32_t)n_slots, lc, d_order[ord], d_rank, n, d_keys[0]);
        size_t tb = temp_bytes;
        BVH_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, tb, d_keys[0], d_keys[1], d_order[ord], d_order[ord ^ 1], (int64_t)n, 0, (int)(rbits + sbits), stream));
        ord ^= 1;
        cur ^= 1;
        n_slots *= 2;
        m_hi = (m_hi + 1) / 2;
    }
    k_bvh_root_box<<<1, 1, 0, stream>>>(d_nodes, d_order[ord], d_boxes);
    BVH_CUDA(cudaGetLastError());
    BVH_CUDA(cudaStreamSynchronize(stream));
    const auto t_built = clk::now();

    guard.keep = true;
    *d_nodes_out = d_nodes;
    *n_nodes_out = total_nodes;
    max_depth = depth;
    if (times) {
        auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        times->ms_pack = ms(t_begin, t_alloc);
        times->ms_h2d = ms(t_alloc, t_h2d);
        times->ms_device = ms(t_h2d, t_built);
        times->ms_d2h = 0.0;
        times->levels = depth;
    }
    return CR_OK;
}

int gpu_fetch_flat_nodes(int device, void* cuda_stream, const void* d_nodes, uint64_t n, std::vector<FlatNode>& out, std::string& err) {
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    BVH_CUDA(cudaSetDevice(device));
    out.resize((size_t)n);
    if (n == 0) return CR_OK;
    BVH_CUDA(cudaMemcpyAsync(static_cast<void*>(out.data()), d_nodes, (size_t)n * sizeof(DevNode), cudaMemcpyDeviceToHost, stream));
    BVH_CUDA(cudaStreamSynchronize(stream));
    return CR_OK;
}

int gpu_flatten_nodes(int device, void* cuda_stream, const void* d_flat, uint64_t n64, void* d_n64, void* d_n32, float* bmax, float* bsmall,
                      std::string& err) {
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    BVH_CUDA(cudaSetDevice(device));
    *bmax = *bsmall = 0.f;
    if (n64 == 0) return CR_OK;
    const uint32_t n = (uint32_t)n64;
    const uint32_t tpb = 256, grid_n = (n + tpb - 1) / tpb;
    DeviceBuffers buf{stream, {}};
    float *d_nb, *d_sorted;
    void* d_temp;
    BVH_CUDA(buf.alloc(&d_nb, n));
    BVH_CUDA(buf.alloc(&d_sorted, n));
    size_t temp_bytes = 0;
    BVH_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, d_nb, d_sorted, (int64_t)n, 0, 32, stream));
    BVH_CUDA(buf.alloc(reinterpret_cast<uint8_t**>(&d_temp), temp_bytes));
    const DevNode* flat = static_cast<const DevNode*>(d_flat);
    k_node_bound<<<grid_n, tpb, 0, stream>>>(flat, n, d_nb);
    BVH_CUDA(cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, d_nb, d_sorted, (int64_t)n, 0, 32, stream));
    // the ~90 % smallest bounds share the tight filter constant, the rest get BIGBOX_BIT (as upload_scene does on the host)
    const size_t k90 = (size_t)((n - 1) * 0.9);
    float picks[2];
    BVH_CUDA(cudaMemcpyAsync(&picks[0], d_sorted + k90, sizeof(float), cudaMemcpyDeviceToHost, stream));
    BVH_CUDA(cudaMemcpyAsync(&picks[1], d_sorted + (n - 1), sizeof(float), cudaMemcpyDeviceToHost, stream));
    BVH_CUDA(cudaStreamSynchronize(stream));
    *bsmall = picks[0];
    *bmax = picks[1];
    k_flatten_nodes<<<grid_n, tpb, 0, stream>>>(flat, d_nb, n, *bsmall, static_cast<NodeRec<double>*>(d_n64), static_cast<NodeRec<float>*>(d_n32));
    BVH_CUDA(cudaGetLastError());
    return CR_OK;
}

int gpu_flatten_tris(int device, void* cuda_stream, const double* h_abc, uint64_t n64, void* d_t64, void* d_t32, std::string& err) {
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    BVH_CUDA(cudaSetDevice(device));
    if (n64 == 0) return CR_OK;
    const uint32_t n = (uint32_t)n64;
    DeviceBuffers buf{stream, {}};
    double* d_abc;
    BVH_CUDA(buf.alloc(&d_abc, (size_t)9 * n));
    BVH_CUDA(StagedCopier::copy(device, d_abc, h_abc, (size_t)9 * n * sizeof(double), stream));
    k_flatten_tris<<<(n + 255) / 256, 256, 0, stream>>>(d_abc, n, static_cast<TriRec<double>*>(d_t64), static_cast<TriRec<float>*>(d_t32));
    BVH_CUDA(cudaGetLastError());
    return CR_OK;
}

}  // namespace crb
