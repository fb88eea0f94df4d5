// bvh_host.h — host-side BVH build types shared by the C-ABI layer (api.cu: the reference-order host builder)
// and the device builder (bvh_build.cu).  Geometry in reference arithmetic (src/objects/bvh.rs, bvhwrapper.rs).
#pragma once
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_pool.h"

namespace crb {

static const double INF = std::numeric_limits<double>::infinity();

// ---- host geometry in reference arithmetic (this file is compiled with -ffp-contract=off) ---------
struct Box {  // Aabb, src/objects/bvh.rs:19-34; default = EMPTY intervals
    double lo[3] = {INF, INF, INF};
    double hi[3] = {-INF, -INF, -INF};
};
// Aabb::new_from_boxes / Interval::tight_enclose, bvh.rs:69-75, utils.rs:631-635
inline Box box_union(const Box& a, const Box& b) {
    Box r;
    for (int k = 0; k < 3; ++k) {
        r.lo[k] = (a.lo[k] <= b.lo[k]) ? a.lo[k] : b.lo[k];
        r.hi[k] = (a.hi[k] >= b.hi[k]) ? a.hi[k] : b.hi[k];
    }
    return r;
}
// Aabb::new_from_points, bvh.rs:46-66
inline Box box_from_points(const double a[3], const double b[3]) {
    Box r;
    for (int k = 0; k < 3; ++k) {
        if (a[k] <= b[k]) {
            r.lo[k] = a[k];
            r.hi[k] = b[k];
        } else {
            r.lo[k] = b[k];
            r.hi[k] = a[k];
        }
    }
    return r;
}
// Aabb::longest_axis, bvh.rs:82-94
inline int longest_axis(const Box& b) {
    const double sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
    if (sx > sy) return (sx > sz) ? 0 : 2;
    if (sy > sz) return 1;
    return 2;
}

struct Element {  // one entry of Scene.elements (scene/mod.rs:77), insertion order
    uint32_t kind, idx;
    bool hide;
    Box box;
};
template <>
struct pool_no_init<Element> : std::true_type {};  // kind / idx / hide are written by the add call, the box by ensure_boxes at commit
typedef HostVec<Element> ElementVec;  // 64 B per primitive: 640 MB for the 10 M-triangle scene, cached between scenes (host_pool.h)

static constexpr uint32_t FLAT_ALWAYS_PASS = 0x100u;  // FlatNode::axis flag: member of a nested HitList, no box test (api.cu GroupEmitter)
static constexpr uint32_t FLAT_GROUP_NODE = 0x200u;   // FlatNode::axis flag: box node whose children are nested elements, not plain nodes

struct FlatNode {
    Box box;
    uint32_t left, right, axis;  // children: node indices, or primitive refs for a leaf node (axis: bits 0..1 + FLAT_* flags)
    uint32_t lchild, rchild;     // node children (REF_NONE for a leaf node)
    uint32_t skip;               // preorder index of the first node after this subtree
};


// number of nodes BVHWrapper::help_generate creates for a span (bvhwrapper.rs:46-80)
inline uint64_t node_count(uint64_t span) {
    if (span <= 2) return 1;
    return 1 + node_count(span / 2) + node_count(span - span / 2);
}

struct BvhBuildTimes {  // wall-clock milliseconds of the phases of one device build
    double ms_pack = 0, ms_h2d = 0, ms_device = 0, ms_d2h = 0;
    uint32_t levels = 0;
};

// Device build of the SAME tree (bvh_build.cu): level-synchronous median split, one stable radix sort per level.
// The FlatNode records stay ON THE DEVICE (*d_nodes, node_count(visible.size()) entries in preorder, root box already
// re-derived as BVHWrapper::new_from_vec does); the caller owns the allocation (cudaFreeAsync on the same stream).
int gpu_build_bvh(int device, void* cuda_stream, const ElementVec& elements, const std::vector<uint32_t>& visible,
                  void** d_nodes, uint64_t* n_nodes, uint32_t& max_depth, BvhBuildTimes* times, std::string& err);
// introspection: copies the device tree to the host
int gpu_fetch_flat_nodes(int device, void* cuda_stream, const void* d_nodes, uint64_t n, std::vector<FlatNode>& out, std::string& err);
// FlatNode (device) -> NodeRec<double> / NodeRec<float> (outward-rounded boxes, BIGBOX bit); returns the two filter bounds
int gpu_flatten_nodes(int device, void* cuda_stream, const void* d_flat, uint64_t n, void* d_n64, void* d_n32, float* bmax, float* bsmall,
                      std::string& err);
// a,b,c (host, [n][9]) -> TriRec<double> / TriRec<float> on the device
int gpu_flatten_tris(int device, void* cuda_stream, const double* h_abc, uint64_t n, void* d_t64, void* d_t32, std::string& err);

}  // namespace crb
