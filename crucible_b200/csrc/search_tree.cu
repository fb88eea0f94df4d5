// search_tree.cu — device build of the SEARCH tree of the order-free trace engine for large scenes (fast_trace.cuh).
//
// The search tree never decides a hit (every candidate is confirmed against its REFERENCE leaf-node box), so it only has
// to be conservative and good.  Scenes below 32 768 primitives get a binned-SAH tree from the host (fast_tree.h); above
// that a linear BVH is built here from what the committed scene already holds on the device — no second upload:
//   1. k_collect : every leaf node of the reference tree (NodeRec<double>, preorder) yields its primitives as
//                  (primitive ref, leaf node << 1 | slot = DFS rank), a conservative f32 box computed from the device
//                  primitive records, and the 63-bit Morton code of the box centre inside the root box;
//   2. cub::DeviceRadixSort by Morton code (library primitive);
//   3. k_hierarchy: Karras' binary radix tree over the sorted codes (ties broken by position);
//   4. k_fit     : bottom-up box fit (second child to arrive folds), with subtree heights;
//   5. k_emit    : 64 B records holding BOTH child boxes; ranges of one or two primitives become leaves.
// Config 4 (10 M triangles): ~25 ms instead of 5.3 s for the host SAH build.
#include <cub/device/device_radix_sort.cuh>

#include <string>

#include "fast_tree.h"
#include "integrator.h"

namespace crb {
namespace {

struct Box32 {
    float lo[3], hi[3];
};

__device__ __forceinline__ float down32(double x) { return __double2float_rd(x); }
__device__ __forceinline__ float up32(double x) { return __double2float_ru(x); }

// conservative box of a primitive from its device record (slightly padded: the records hold e1 = b - a, not b)
__device__ Box32 prim_box32(uint32_t ref, const SphereRec<double>* sph, const TriRec<double>* tri, const QuadRec<double>* quad) {
    const uint32_t kind = ref_kind(ref), idx = ref_index(ref);
    double lo[3], hi[3];
    if (kind == CR_PRIM_SPHERE) {
        const SphereRec<double> s = sph[idx];
        const double c[3] = {s.cx, s.cy, s.cz};
        for (int k = 0; k < 3; ++k) {
            lo[k] = c[k] - s.r;
            hi[k] = c[k] + s.r;
        }
    } else if (kind == CR_PRIM_TRIANGLE) {
        const TriRec<double> t = tri[idx];
        const double a[3] = {t.ax, t.ay, t.az}, e1[3] = {t.e1x, t.e1y, t.e1z}, e2[3] = {t.e2x, t.e2y, t.e2z};
        for (int k = 0; k < 3; ++k) {
            const double b = a[k] + e1[k], c = a[k] + e2[k];
            lo[k] = fmin(a[k], fmin(b, c));
            hi[k] = fmax(a[k], fmax(b, c));
        }
    } else {
        const QuadRec<double> q = quad[idx];
        const double Q[3] = {q.qx, q.qy, q.qz}, u[3] = {q.ux, q.uy, q.uz}, v[3] = {q.vx, q.vy, q.vz};
        for (int k = 0; k < 3; ++k) {
            const double p1 = Q[k] + u[k], p2 = Q[k] + v[k], p3 = p1 + v[k];
            lo[k] = fmin(fmin(Q[k], p1), fmin(p2, p3)) - 5.0e-5;  // the reference pads flat quad boxes by 1e-4 / 2
            hi[k] = fmax(fmax(Q[k], p1), fmax(p2, p3)) + 5.0e-5;
        }
    }
    Box32 b;
    for (int k = 0; k < 3; ++k) {
        const double pad = 1.0e-12 * (fabs(lo[k]) + fabs(hi[k])) + 1.0e-30;
        b.lo[k] = down32(lo[k] - pad);
        b.hi[k] = up32(hi[k] + pad);
    }
    return b;
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {  // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_collect(const NodeRec<double>* __restrict__ nodes, uint32_t n_nodes, const SphereRec<double>* sph, const TriRec<double>* tri,
                          const QuadRec<double>* quad, uint32_t* __restrict__ counter, uint2* __restrict__ entries, Box32* __restrict__ boxes,
                          uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const NodeRec<double> root = nodes[0];
    const double r_lo[3] = {root.xmin, root.ymin, root.zmin};
    const double r_ext[3] = {root.xmax - root.xmin, root.ymax - root.ymin, root.zmax - root.zmin};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const uint32_t wa = nodes[i].left, wb = nodes[i].right;
        if (!ref_is_leaf(wa)) continue;
        const uint32_t cnt = wb != REF_NONE ? 2u : 1u;
        const uint32_t pos = atomicAdd(counter, cnt);
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t ref = k == 0 ? (wa & ~(1u << 27)) : wb;  // bit 27 of the first word is the node's BIGBOX flag
            const Box32 b = prim_box32(ref, sph, tri, quad);
            entries[pos + k] = make_uint2(ref, 2u * i + k);
            boxes[pos + k] = b;
            uint64_t code = 0;
            for (int a = 0; a < 3; ++a) {
                double t = r_ext[a] > 0.0 ? (0.5 * ((double)b.lo[a] + (double)b.hi[a]) - r_lo[a]) / r_ext[a] : 0.0;
                t = fmin(fmax(t, 0.0), 1.0);
                code |= spread21((uint64_t)(t * 2097151.0)) << a;
            }
            keys[pos + k] = code;
            vals[pos + k] = pos + k;
        }
    }
}

// common-prefix length of the sorted codes at positions i and j (Karras 2012); ties broken by position
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, uint32_t n, int i, int j) {
    if (j < 0 || j >= (int)n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz((uint32_t)i ^ (uint32_t)j) : __clzll((long long)(a ^ b));
}

struct Radix {
    uint32_t left, right;  // child: internal index, or leaf position | 0x80000000
    uint32_t first, last;  // covered range of sorted positions
};

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, uint32_t n, Radix* __restrict__ nodes, uint32_t* __restrict__ parent_int,
                            uint32_t* __restrict__ parent_leaf) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)n - 1; i += gridDim.x * blockDim.x) {
        const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(keys, n, i, i - d);
        int lmax = 2;
        while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
        int l = 0;
        for (int t = lmax / 2; t >= 1; t /= 2)
            if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
        const int j = i + l * d;
        const int dnode = delta(keys, n, i, j);
        int s = 0;
        for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
            if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
            if (t == 1) break;
        }
        const int gamma = i + s * d + min(d, 0);
        const int first = min(i, j), last = max(i, j);
        Radix r;
        r.first = (uint32_t)first;
        r.last = (uint32_t)last;
        if (first == gamma) {
            r.left = (uint32_t)gamma | 0x80000000u;
            parent_leaf[gamma] = (uint32_t)i;
        } else {
            r.left = (uint32_t)gamma;
            parent_int[gamma] = (uint32_t)i;
        }
        if (last == gamma + 1) {
            r.right = (uint32_t)(gamma + 1) | 0x80000000u;
            parent_leaf[gamma + 1] = (uint32_t)i;
        } else {
            r.right = (uint32_t)(gamma + 1);
            parent_int[gamma + 1] = (uint32_t)i;
        }
        nodes[i] = r;
        if (i == 0) parent_int[0] = 0xFFFFFFFFu;
    }
}

__global__ void k_fit(const Radix* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ parent_int, const uint32_t* __restrict__ parent_leaf,
                      const uint32_t* __restrict__ vals, const Box32* __restrict__ leaf_boxes, Box32* __restrict__ node_boxes,
                      uint32_t* __restrict__ height, uint32_t* __restrict__ arrived) {
    for (uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x; leaf < n; leaf += gridDim.x * blockDim.x) {
        uint32_t cur = parent_leaf[leaf];
        while (cur != 0xFFFFFFFFu) {
            __threadfence();
            if (atomicAdd(&arrived[cur], 1u) == 0u) break;  // the first child to arrive stops; the second folds
            __threadfence();
            const Radix r = nodes[cur];
            Box32 b;
            uint32_t h = 0;
            for (int side = 0; side < 2; ++side) {
                const uint32_t c = side == 0 ? r.left : r.right;
                Box32 cb;
                uint32_t ch = 0;
                if (c & 0x80000000u) {
                    cb = leaf_boxes[vals[c & 0x7FFFFFFFu]];
                } else {  // written by another thread before it passed this node's counter: read around L1
                    const volatile float* vb = reinterpret_cast<const volatile float*>(&node_boxes[c]);
                    for (int k = 0; k < 3; ++k) {
                        cb.lo[k] = vb[k];
                        cb.hi[k] = vb[3 + k];
                    }
                    ch = *reinterpret_cast<const volatile uint32_t*>(&height[c]);
                }
                if (side == 0) {
                    b = cb;
                } else {
                    for (int k = 0; k < 3; ++k) {
                        b.lo[k] = fminf(b.lo[k], cb.lo[k]);
                        b.hi[k] = fmaxf(b.hi[k], cb.hi[k]);
                    }
                }
                h = max(h, ch);
            }
            node_boxes[cur] = b;
            height[cur] = h + 1u;
            cur = parent_int[cur];
        }
    }
}

__global__ void k_emit(const Radix* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ vals, const Box32* __restrict__ leaf_boxes,
                       const Box32* __restrict__ node_boxes, const uint2* __restrict__ entries, FastNodeRec* __restrict__ out,
                       uint2* __restrict__ table) {
    const uint32_t total = 2u * n - 1u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        if (i >= n - 1u) {  // leaf-primitive table in sorted order
            const uint32_t pos = i - (n - 1u);
            table[pos] = entries[vals[pos]];
            continue;
        }
        const Radix r = nodes[i];
        if (r.last - r.first + 1u < 3u) continue;  // folded into a leaf of its parent
        FastNodeRec rec;
        for (int side = 0; side < 2; ++side) {
            const uint32_t c = side == 0 ? r.left : r.right;
            Box32 b;
            uint32_t word;
            if (c & 0x80000000u) {
                const uint32_t pos = c & 0x7FFFFFFFu;
                b = leaf_boxes[vals[pos]];
                word = FAST_LEAF | pos;
            } else {
                const Radix cr = nodes[c];
                b = node_boxes[c];
                word = (cr.last - cr.first + 1u <= 2u) ? (FAST_LEAF | ((cr.last - cr.first) << 28) | cr.first) : c;
            }
            FastHalf& h = rec.c[side];
            h.xmin = b.lo[0]; h.xmax = b.hi[0];
            h.ymin = b.lo[1]; h.ymax = b.hi[1];
            h.zmin = b.lo[2]; h.zmax = b.hi[2];
            h.child = word;
            h.pad = 0;
        }
        out[i] = rec;
    }
}

#define ST_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            err = std::string("search tree build: " #call ": ") + cudaGetErrorString(e__); \
            cleanup();                                                                 \
            return CR_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

}  // namespace

// Builds the search tree of a committed scene (>= 3 visible primitives) from its device records.  On success
// *d_fast_nodes / *d_fast_prims are stream-ordered allocations the caller owns; a tree deeper than the traversal stack
// (FAST_MAX_DEPTH) is discarded and both come back nullptr (the scene then keeps reference order).
int gpu_build_search_tree(const SceneDeviceData& d, cudaStream_t stream, uint32_t n_visible, void** d_fast_nodes, void** d_fast_prims,
                          uint32_t* depth_out, std::string& err) {
    *d_fast_nodes = *d_fast_prims = nullptr;
    *depth_out = 0;
    const uint32_t n = n_visible;
    if (n < 3 || d.n_nodes == 0) return CR_OK;
    std::vector<void*> tmp;
    void* keep[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        for (void* p : tmp) cudaFreeAsync(p, stream);
        for (void* p : keep)
            if (p) cudaFreeAsync(p, stream);
    };
    auto alloc = [&](void** p, size_t bytes, bool temporary) {
        const cudaError_t e = cudaMallocAsync(p, bytes, stream);
        if (e == cudaSuccess && temporary) tmp.push_back(*p);
        return e;
    };
    uint32_t *counter, *vals[2], *parent_int, *parent_leaf, *height, *arrived;
    uint2 *entries, *table;
    Box32 *leaf_boxes, *node_boxes;
    uint64_t* keys[2];
    Radix* radix;
    FastNodeRec* out;
    void* cub_tmp;
    ST_CUDA(alloc((void**)&counter, 256, true));
    ST_CUDA(alloc((void**)&entries, (size_t)n * sizeof(uint2), true));
    ST_CUDA(alloc((void**)&leaf_boxes, (size_t)n * sizeof(Box32), true));
    ST_CUDA(alloc((void**)&node_boxes, (size_t)n * sizeof(Box32), true));
    ST_CUDA(alloc((void**)&keys[0], (size_t)n * sizeof(uint64_t), true));
    ST_CUDA(alloc((void**)&keys[1], (size_t)n * sizeof(uint64_t), true));
    ST_CUDA(alloc((void**)&vals[0], (size_t)n * sizeof(uint32_t), true));
    ST_CUDA(alloc((void**)&vals[1], (size_t)n * sizeof(uint32_t), true));
    ST_CUDA(alloc((void**)&parent_int, (size_t)n * sizeof(uint32_t), true));
    ST_CUDA(alloc((void**)&parent_leaf, (size_t)n * sizeof(uint32_t), true));
    ST_CUDA(alloc((void**)&height, (size_t)n * sizeof(uint32_t), true));
    ST_CUDA(alloc((void**)&arrived, (size_t)n * sizeof(uint32_t), true));
    ST_CUDA(alloc((void**)&radix, (size_t)n * sizeof(Radix), true));
    ST_CUDA(alloc((void**)&out, (size_t)(n - 1) * sizeof(FastNodeRec), false));
    keep[0] = out;
    ST_CUDA(alloc((void**)&table, (size_t)n * sizeof(uint2), false));
    keep[1] = table;
    size_t cub_bytes = 0;
    ST_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys[0], keys[1], vals[0], vals[1], (int64_t)n, 0, 63, stream));
    ST_CUDA(alloc(&cub_tmp, cub_bytes, true));
    ST_CUDA(cudaMemsetAsync(counter, 0, 256, stream));
    ST_CUDA(cudaMemsetAsync(arrived, 0, (size_t)n * sizeof(uint32_t), stream));
    const int tpb = 256;
    auto grid = [&](size_t items) { return (int)std::min<size_t>((items + tpb - 1) / tpb, (size_t)d.num_sms * 32); };
    k_collect<<<grid(d.n_nodes), tpb, 0, stream>>>(static_cast<const NodeRec<double>*>(d.nodes[0]), d.n_nodes,
                                                   static_cast<const SphereRec<double>*>(d.spheres[0]), static_cast<const TriRec<double>*>(d.tris[0]),
                                                   static_cast<const QuadRec<double>*>(d.quads[0]), counter, entries, leaf_boxes, keys[0], vals[0]);
    ST_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys[0], keys[1], vals[0], vals[1], (int64_t)n, 0, 63, stream));
    k_hierarchy<<<grid(n), tpb, 0, stream>>>(keys[1], n, radix, parent_int, parent_leaf);
    k_fit<<<grid(n), tpb, 0, stream>>>(radix, n, parent_int, parent_leaf, vals[1], leaf_boxes, node_boxes, height, arrived);
    k_emit<<<grid(2 * (size_t)n), tpb, 0, stream>>>(radix, n, vals[1], leaf_boxes, node_boxes, entries, out, table);
    ST_CUDA(cudaGetLastError());
    uint32_t h_info[2] = {0, 0};  // primitives collected, height of the root
    ST_CUDA(cudaMemcpyAsync(&h_info[0], counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    ST_CUDA(cudaMemcpyAsync(&h_info[1], height, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    ST_CUDA(cudaStreamSynchronize(stream));
    for (void* p : tmp) cudaFreeAsync(p, stream);
    tmp.clear();
    if (h_info[0] != n) {
        err = "search tree build: the reference tree holds " + std::to_string(h_info[0]) + " primitives, expected " + std::to_string(n);
        cleanup();
        return CR_ERR_STATE;
    }
    if (h_info[1] + 1u > (uint32_t)FAST_MAX_DEPTH) {  // deeper than the traversal stack: keep reference order for this scene
        cleanup();
        return CR_OK;
    }
    *d_fast_nodes = out;
    *d_fast_prims = table;
    *depth_out = h_info[1] + 1u;
    return CR_OK;
}

}  // namespace crb
