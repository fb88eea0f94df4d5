// api.cu — the C ABI of include/crucible_gpu.h: scene staging, the host BVH build that reproduces
// BVHWrapper::new_wrapper (src/objects/bvhwrapper.rs:15-94) and flattens it to preorder records,
// device upload, and the trace / render entry points.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <mutex>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

#include "bvh_host.h"
#include "fast_tree.h"
#include "h2d_staged.h"
#include "integrator.h"

using namespace crb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define API_CUDA(call)                                                                                     \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) return fail(CR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct HostImage {
    int w, h;
    HostVec<uint8_t> rgb;  // a 4096 x 2048 sky is 25 MB: cached block (host_pool.h)
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
};

// CUDA arrays (+ their point-sampled texture objects) of destroyed scenes, per device, by size.  A caller that rebuilds its
// scene every frame (Scene::render_image, scene/mod.rs:332-347) would otherwise pay cudaMallocArray / cudaFreeArray for a 32 MB
// sky image per frame; measured on the teapot config, cudaFreeArray inside cr_scene_destroy took 0.6 ... 256 ms (it unmaps
// device memory and synchronises the device) and made the end-to-end rate jump between 1270 and 2040 Msamples/s from run to
// run.  Only the ALLOCATION is reused: every scene uploads its texels again.  cr_device_trim frees the cache.
struct CachedArray {
    int w, h;
    cudaArray_t arr;
    cudaTextureObject_t tex;
};
struct ImageArrayCache {
    static constexpr size_t MAX_BYTES = (size_t)1 << 30;
    std::mutex mu;
    std::vector<CachedArray> free_list;
    size_t bytes = 0;
    bool take(int w, int h, cudaArray_t* arr, cudaTextureObject_t* tex) {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < free_list.size(); ++i)
            if (free_list[i].w == w && free_list[i].h == h) {
                *arr = free_list[i].arr;
                *tex = free_list[i].tex;
                bytes -= (size_t)w * h * 4;
                free_list.erase(free_list.begin() + (long)i);
                return true;
            }
        return false;
    }
    // the caller has made sure no work that reads the array is still in flight
    void give(int w, int h, cudaArray_t arr, cudaTextureObject_t tex) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (bytes + (size_t)w * h * 4 <= MAX_BYTES) {
                free_list.push_back(CachedArray{w, h, arr, tex});
                bytes += (size_t)w * h * 4;
                return;
            }
        }
        if (tex) cudaDestroyTextureObject(tex);
        if (arr) cudaFreeArray(arr);
    }
    void trim() {
        std::vector<CachedArray> gone;
        {
            std::lock_guard<std::mutex> lk(mu);
            gone.swap(free_list);
            bytes = 0;
        }
        for (auto& c : gone) {
            if (c.tex) cudaDestroyTextureObject(c.tex);
            if (c.arr) cudaFreeArray(c.arr);
        }
    }
};
ImageArrayCache& image_cache(int device) {
    static ImageArrayCache* c = new ImageArrayCache[64];  // never destroyed (CUDA objects must not be freed after the runtime)
    return c[device < 0 || device >= 64 ? 0 : device];
}

}  // namespace

struct CrScene {
    int device = -1;  // -1 = host-only (no CUDA device): build/introspection work, compute fails
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    // staging
    ElementVec elements;  // the big staging arrays are HostVec: their blocks are cached between scenes (host_pool.h)
    size_t boxed = 0;  // elements[0, boxed) have their box; the rest get it at commit (ensure_boxes: all host threads at once)
    HostVec<double> spheres;  // [n][4]
    HostVec<double> tris;     // [n][9] a,b,c
    HostVec<double> quads;    // [n][9] Q,u,v
    HostVec<int32_t> mat_of[3], obj_of[3], prim_of[3];
    std::vector<CrMaterial> mats;
    std::vector<CrTexture> texs;
    std::vector<HostImage> images;
    int sky_kind = CR_SKY_DEFAULT, sky_image = -1;
    std::map<uint64_t, std::vector<CrAnimKey>> anim;  // (prim_index * 4 + point) -> keyframes in timeline order
    // nested elements (Scene::add_element accepts Hittables::HitList / BVHWrapper, scene/mod.rs:160-166).  A member is an
    // element index, or GROUP_MEMBER | group id.  `top` (the top-level member list) exists once the scene has a group.
    struct Group {
        int32_t kind, parent;
        std::vector<uint32_t> members;
    };
    std::vector<Group> groups;
    std::vector<int32_t> open_groups;
    std::vector<uint32_t> top;
    // built
    bool committed = false;
    std::vector<FlatNode> nodes;
    uint32_t root = REF_MISS;
    uint32_t max_depth = 0;
    uint64_t n_visible = 0;
    std::vector<int32_t> leaf_order_grouped;  // scenes with groups: primitives in visiting order (GroupEmitter)
    int bvh_builder = CR_BVH_AUTO;
    CrCommitInfo commit_info = {};
    // device-built tree: FlatNode records on the device; `nodes` is then filled on demand (host_nodes)
    void* d_flat_nodes = nullptr;
    uint64_t n_nodes = 0;
    // device
    SceneDeviceData dev;
    std::vector<void*> dev_allocs;
    void* d_out_rgb = nullptr;
    void* d_out_rgb8 = nullptr;
    size_t out_cap = 0;
    void* d_io = nullptr;  // trace_batch staging
    size_t io_cap = 0;
    int64_t last_retried = 0;  // rays of the last cr_trace_batch the order-free engine handed back
    // cr_render_multi: the assembled image on this (the first) device; plain cudaMalloc so that peers can map it
    void* d_multi_rgb = nullptr;
    void* d_multi_rgb8 = nullptr;
    size_t multi_cap = 0;

    void free_flat_tree() {
        if (d_flat_nodes) cudaFreeAsync(d_flat_nodes, stream);
        d_flat_nodes = nullptr;
    }
    void free_device_scene() {
        for (void* p : dev_allocs) cudaFreeAsync(p, stream);
        dev_allocs.clear();
        bool drained = false;
        for (auto& im : images) {
            if (im.arr) {
                // a render of this scene may still be in flight on a caller's stream (cr_render_device is asynchronous);
                // cudaFreeArray used to wait for it implicitly
                if (!drained) cudaDeviceSynchronize();
                drained = true;
                image_cache(device).give(im.w, im.h, im.arr, im.tex);
            } else if (im.tex) {
                cudaDestroyTextureObject(im.tex);
            }
            im.tex = 0;
            im.arr = nullptr;
        }
        dev = SceneDeviceData();
    }
};

namespace {

// One grow-only scratch arena per device, shared by every scene of the process (path pool, queues,
// fixed-point framebuffer): a second render, or a second scene, reuses the memory instead of paying
// cudaMalloc for gigabytes again.  `mu` serialises the renders of one device (the arena is shared) and guards
// cr_device_trim; different devices render concurrently (cr_render_multi drives one host thread per device).
struct DeviceSlot {
    std::mutex mu;
    Workspace ws;
};
DeviceSlot& device_slot(int device) {
    static DeviceSlot slots[64];
    return slots[device < 0 || device >= 64 ? 0 : device];
}
Workspace& device_workspace(int device) { return device_slot(device).ws; }

// Per-device facts, queried once per process: cudaGetDeviceProperties costs 3-50 ms per call on an 8-GPU box
// (measured: it dominated the end-to-end step when every frame created its scene, as Scene::render_image does).
struct DeviceInfo {
    int state = 0;  // 0 = not queried, 1 = usable sm_100 device, -1 = unusable
    int num_sms = 0;
    std::string why;
};
DeviceInfo& device_info(int device) {
    static DeviceInfo info[64];
    static DeviceInfo bad;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);  // first calls may come from several host threads
    if (device < 0 || device >= 64) {
        bad.state = -1;
        bad.why = "CUDA device " + std::to_string(device) + " not available";
        return bad;
    }
    DeviceInfo& d = info[device];
    if (d.state != 0) return d;
    int n = 0;
    cudaDeviceProp p;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n) {
        cudaGetLastError();
        d.why = "CUDA device " + std::to_string(device) + " not available";
        d.state = -1;
    } else if (cudaGetDeviceProperties(&p, device) != cudaSuccess || p.major != 10) {
        cudaGetLastError();
        d.why = "device is not sm_100 (this library ships sm_100a code only)";
        d.state = -1;
    } else {
        d.num_sms = p.multiProcessorCount;
        // scene buffers come from the device's stream-ordered pool and stay cached there between scenes:
        // a per-frame scene rebuild then costs no cudaMalloc / cudaFree (each of which synchronises the device).
        // cr_device_trim() gives the memory back and restores the default threshold.
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        d.state = 1;  // published last: the fields above are complete when another thread sees it
    }
    return d;
}


struct Builder {
    CrScene& sc;
    std::vector<uint32_t>& objs;  // indices into sc.elements (the cloned, re-sorted list)
    std::vector<FlatNode>& nodes;
    uint32_t max_depth = 0;

    uint32_t leaf_ref(uint32_t e) const { return make_leaf(sc.elements[e].kind, sc.elements[e].idx); }

    // Writes the subtree of objs[start,end) at nodes[at...] in preorder and returns its depth.
    uint32_t build(size_t start, size_t end, uint32_t at, int par_depth) {
        Box bbox;  // Aabb::default()
        for (size_t i = start; i < end; ++i) bbox = box_union(bbox, sc.elements[objs[i]].box);
        const int axis = longest_axis(bbox);
        const size_t span = end - start;
        FlatNode n;
        n.box = bbox;
        n.axis = (uint32_t)axis;
        uint32_t depth = 1;
        n.skip = at + (uint32_t)node_count(span);
        n.lchild = n.rchild = REF_NONE;
        if (span == 1) {
            n.left = leaf_ref(objs[start]);
            n.right = REF_NONE;  // reference stores the same object twice; second test provably None
        } else if (span == 2) {
            n.left = leaf_ref(objs[start]);
            n.right = leaf_ref(objs[start + 1]);
        } else {
            // stable sort by bbox.min on the axis (box_compare, bvhwrapper.rs:82-94; NaN -> Equal)
            struct Key {
                double k;
                uint32_t e;
            };
            std::vector<Key> keys(span);
            for (size_t i = 0; i < span; ++i) keys[i] = {sc.elements[objs[start + i]].box.lo[axis], objs[start + i]};
            std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.k < b.k; });
            for (size_t i = 0; i < span; ++i) objs[start + i] = keys[i].e;
            keys.clear();
            keys.shrink_to_fit();
            const size_t mid = start + span / 2;
            const uint32_t l_at = at + 1;
            const uint32_t r_at = at + 1 + (uint32_t)node_count(mid - start);
            n.left = n.lchild = l_at;
            n.right = n.rchild = r_at;
            uint32_t dl = 0, dr = 0;
            if (par_depth > 0 && span > 65536) {
                std::thread th([&] { dl = build(start, mid, l_at, par_depth - 1); });
                dr = build(mid, end, r_at, par_depth - 1);
                th.join();
            } else {
                dl = build(start, mid, l_at, 0);
                dr = build(mid, end, r_at, 0);
            }
            depth = 1 + std::max(dl, dr);
        }
        nodes[at] = n;
        return depth;
    }
};

// ---- scenes with nested elements ---------------------------------------------------------------------------------------
// The reference's world is a tree of Hittables: BVHWrapper nodes (box test, left, right: bvhwrapper.rs:97-126), HitLists
// (linear scan with ONE shrinking interval and no box test: hitlist.rs:52-65) and primitives.  A nested BVHWrapper met as
// a leaf of its parent's tree tests its own box and recurses, i.e. it behaves as if its subtree were grafted into the
// parent; a nested HitList is a run of members visited in order.  Both flatten into the SAME preorder array the stackless
// walk already uses: box nodes with a skip link, and "always pass" leaf nodes (FLAT_ALWAYS_PASS) holding one or two
// consecutive list members, which the walk enters without a box test.  A span-1 node stores its element twice
// (bvhwrapper.rs:57-59); the second visit runs with the interval closed at the first visit's hit and can accept nothing
// (every primitive it reaches was reached, with a wider interval, by the first), so the element is emitted once.
static constexpr uint32_t GROUP_MEMBER = 0x80000000u;
struct GroupEmitter {
    CrScene& sc;
    std::vector<FlatNode>& out;
    std::vector<Box> gbox;           // box of every group as its parent sees it
    std::vector<int32_t> leaf_order; // primitives in visiting order
    uint32_t max_depth = 0;

    bool is_group(uint32_t m) const { return (m & GROUP_MEMBER) != 0; }
    const Box& box_of(uint32_t m) const { return is_group(m) ? gbox[m & ~GROUP_MEMBER] : sc.elements[m].box; }
    uint32_t leaf_ref(uint32_t e) const { return make_leaf(sc.elements[e].kind, sc.elements[e].idx); }
    // BVHWrapper::new_wrapper keeps nested lists / wrappers and the primitives that are not hidden (bvhwrapper.rs:16-27)
    std::vector<uint32_t> visible_of(const std::vector<uint32_t>& members) const {
        std::vector<uint32_t> v;
        for (uint32_t m : members)
            if (is_group(m) || !sc.elements[m].hide) v.push_back(m);
        return v;
    }
    // Boxes, innermost groups first (a nested group always has a larger id than its parent).  HitList::add folds the
    // members' boxes in insertion order, hidden ones included (hitlist.rs:27-30); a wrapper's box is its root's, the union
    // of its visible members (bvhwrapper.rs:38-41, 47-50; min / max are exact, so the grouping of the fold does not matter).
    void compute_boxes() {
        gbox.assign(sc.groups.size(), Box());
        for (size_t g = sc.groups.size(); g-- > 0;) {
            const CrScene::Group& grp = sc.groups[g];
            Box b;
            for (uint32_t m : grp.members) {
                if (grp.kind == CR_GROUP_BVH && !is_group(m) && sc.elements[m].hide) continue;
                b = box_union(b, box_of(m));
            }
            gbox[g] = b;
        }
    }
    void note_depth(uint32_t d) { max_depth = std::max(max_depth, d); }
    void emit_always_pass(uint32_t e0, uint32_t e1, uint32_t depth) {
        FlatNode n;
        n.box = sc.elements[e0].box;
        if (e1 != REF_NONE) n.box = box_union(n.box, sc.elements[e1].box);  // informational: never tested
        n.left = leaf_ref(e0);
        n.right = e1 == REF_NONE ? REF_NONE : leaf_ref(e1);
        n.axis = FLAT_ALWAYS_PASS;
        n.lchild = n.rchild = REF_NONE;
        n.skip = (uint32_t)out.size() + 1u;
        out.push_back(n);
        leaf_order.push_back(sc.prim_of[sc.elements[e0].kind][sc.elements[e0].idx]);
        if (e1 != REF_NONE) leaf_order.push_back(sc.prim_of[sc.elements[e1].kind][sc.elements[e1].idx]);
        note_depth(depth);
    }
    // one element met without a box of its own: a member of a HitList, or a child of a box node that is not a plain node
    void emit_item(uint32_t m, uint32_t depth) {
        if (!is_group(m)) {
            if (!sc.elements[m].hide) emit_always_pass(m, REF_NONE, depth);  // Sphere::hit / Triangle::hit return None when hidden
            return;
        }
        const CrScene::Group& grp = sc.groups[m & ~GROUP_MEMBER];
        if (grp.kind == CR_GROUP_BVH) {
            std::vector<uint32_t> vis = visible_of(grp.members);
            if (!vis.empty()) emit_bvh(vis, 0, vis.size(), depth);  // empty: the empty HitList of bvhwrapper.rs:29-31
            return;
        }
        // HitList::hit: members in order, one shrinking interval; two consecutive primitives share a leaf node
        for (size_t i = 0; i < grp.members.size(); ++i) {
            const uint32_t a = grp.members[i];
            if (is_group(a)) {
                emit_item(a, depth + 1);
                continue;
            }
            if (sc.elements[a].hide) continue;
            uint32_t b = REF_NONE;
            size_t j = i + 1;
            while (j < grp.members.size() && !is_group(grp.members[j]) && sc.elements[grp.members[j]].hide) ++j;
            if (j < grp.members.size() && !is_group(grp.members[j])) {
                b = grp.members[j];
                i = j;
            }
            emit_always_pass(a, b, depth + 1);
        }
    }
    // BVHWrapper::help_generate (bvhwrapper.rs:46-80) over elements that may be nested
    void emit_bvh(std::vector<uint32_t>& items, size_t start, size_t end, uint32_t depth) {
        Box bbox;
        for (size_t i = start; i < end; ++i) bbox = box_union(bbox, box_of(items[i]));
        const int axis = longest_axis(bbox);
        const size_t span = end - start;
        const uint32_t at = (uint32_t)out.size();
        FlatNode n;
        n.box = bbox;
        n.axis = (uint32_t)axis;
        n.lchild = n.rchild = REF_NONE;
        n.left = n.right = REF_NONE;
        out.push_back(n);
        note_depth(depth);
        if (span <= 2) {
            const uint32_t a = items[start], b = span == 2 ? items[start + 1] : REF_NONE;
            if (!is_group(a) && (b == REF_NONE || !is_group(b))) {  // the plain leaf node
                n.left = leaf_ref(a);
                n.right = b == REF_NONE ? REF_NONE : leaf_ref(b);
                leaf_order.push_back(sc.prim_of[sc.elements[a].kind][sc.elements[a].idx]);
                if (b != REF_NONE) leaf_order.push_back(sc.prim_of[sc.elements[b].kind][sc.elements[b].idx]);
            } else {
                n.axis |= FLAT_GROUP_NODE;
                n.left = n.lchild = at + 1u;
                emit_item(a, depth + 1);
                if (b != REF_NONE) {
                    n.right = n.rchild = (uint32_t)out.size();
                    emit_item(b, depth + 1);
                }
            }
        } else {
            struct Key {
                double k;
                uint32_t m;
            };
            std::vector<Key> keys(span);
            for (size_t i = 0; i < span; ++i) keys[i] = {box_of(items[start + i]).lo[axis], items[start + i]};
            std::stable_sort(keys.begin(), keys.end(), [](const Key& x, const Key& y) { return x.k < y.k; });  // box_compare, :82-94
            for (size_t i = 0; i < span; ++i) items[start + i] = keys[i].m;
            const size_t mid = start + span / 2;
            n.left = n.lchild = at + 1u;
            emit_bvh(items, start, mid, depth + 1);
            n.right = n.rchild = (uint32_t)out.size();
            emit_bvh(items, mid, end, depth + 1);
        }
        n.skip = (uint32_t)out.size();
        out[at] = n;
    }
};

inline float f32_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
inline float f32_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

// fn(a, b) over [0, n) in contiguous ranges, on up to 16 host threads when n is large (staging passes over millions of
// primitives); small n runs on the caller's thread.
template <typename F>
void parallel_ranges(size_t n, F fn) {
    const size_t n_threads = n >= (1u << 16) ? std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (n_threads <= 1) {
        fn((size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (size_t t = 1; t < n_threads; ++t) pool.emplace_back(fn, n * t / n_threads, n * (t + 1) / n_threads);
    fn((size_t)0, n / n_threads);
    for (auto& th : pool) th.join();
}

// the same over n ITEMS of very different sizes (batches): threads when the items carry at least 64 K primitives in total
template <typename F>
void parallel_ranges_always(size_t n_items, size_t weight, F fn) {
    const size_t n_threads = (weight >= (1u << 16) && n_items > 1) ? std::min<size_t>(n_items, std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency()))) : 1;
    if (n_threads <= 1) {
        fn((size_t)0, n_items);
        return;
    }
    std::vector<std::thread> pool;
    for (size_t t = 1; t < n_threads; ++t) pool.emplace_back(fn, n_items * t / n_threads, n_items * (t + 1) / n_threads);
    fn((size_t)0, n_items / n_threads);
    for (auto& th : pool) th.join();
}

template <typename T, typename A>
int upload(CrScene* s, const std::vector<T, A>& host, void** out) {
    *out = nullptr;
    if (host.empty()) return CR_OK;
    void* d = nullptr;
    API_CUDA(cudaMallocAsync(&d, host.size() * sizeof(T), s->stream));
    s->dev_allocs.push_back(d);
    // pageable source: the call returns once the bytes are staged, so `host` may die with the caller's scope
    API_CUDA(StagedCopier::copy(s->device, d, host.data(), host.size() * sizeof(T), s->stream));  // big arrays: several host threads
    *out = d;
    return CR_OK;
}

// does the texture tree under `tex` reach an image texture (the only consumer of u, v)?
bool tex_needs_uv(const CrScene* s, int tex, int depth = 0) {
    if (tex < 0 || tex >= (int)s->texs.size() || depth > MAX_TEX_NEST) return true;
    const CrTexture& t = s->texs[(size_t)tex];
    if (t.kind == CR_TEX_IMAGE) return true;
    if (t.kind == CR_TEX_CHECKER) return tex_needs_uv(s, t.even, depth + 1) || tex_needs_uv(s, t.odd, depth + 1);
    return false;
}

// ---- search tree of the order-free engine: per primitive, its REFERENCE leaf node and slot (= DFS rank) ----------
// A leaf node of the committed reference tree holds one or two primitive refs in its two words (common.cuh); the map is
// read off the device copy, so it serves host-built and device-built trees alike.
__device__ __forceinline__ uint32_t prim_slot(uint32_t ref, uint32_t off_tri, uint32_t off_quad) {
    const uint32_t k = ref_kind(ref);
    return (k == CR_PRIM_SPHERE ? 0u : (k == CR_PRIM_TRIANGLE ? off_tri : off_quad)) + ref_index(ref);
}
__global__ void k_fast_leafmap(const NodeRec<double>* __restrict__ nodes, uint32_t n_nodes, uint32_t* __restrict__ map, uint32_t off_tri,
                               uint32_t off_quad) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const uint32_t wa = nodes[i].left, wb = nodes[i].right;
        if (!ref_is_leaf(wa)) continue;
        map[prim_slot(wa, off_tri, off_quad)] = 2u * i;
        if (wb != REF_NONE) map[prim_slot(wb, off_tri, off_quad)] = 2u * i + 1u;
    }
}
__global__ void k_fast_fill(uint2* __restrict__ prims, uint32_t n, const uint32_t* __restrict__ map, uint32_t off_tri, uint32_t off_quad) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        prims[j].y = map[prim_slot(prims[j].x, off_tri, off_quad)];
}

// Builds the search tree over the visible primitives on the host (fast_tree.h), uploads it and fills the DFS ranks.
int upload_search_tree(CrScene* s, const std::vector<uint32_t>& visible) {
    SceneDeviceData& d = s->dev;
    d.fast_nodes = d.fast_prims = nullptr;
    d.n_fast_nodes = d.n_fast_prims = 0;
    if (visible.empty() || !s->anim.empty() || !s->groups.empty() || getenv("CRB_NO_SEARCH_TREE")) return CR_OK;
    // large scenes: linear BVH built on the device from the records just uploaded (search_tree.cu); small ones: binned SAH here
    bool on_device = visible.size() >= 32768;
    if (const char* e = getenv("CRB_SEARCH_TREE")) on_device = e[0] == 'l' && visible.size() >= 3;
    if (on_device) {
        void *dn = nullptr, *dp = nullptr;
        uint32_t depth = 0;
        std::string err;
        const int rc = gpu_build_search_tree(d, s->stream, (uint32_t)visible.size(), &dn, &dp, &depth, err);
        if (rc != CR_OK) return fail(rc, err);
        if (dn) s->dev_allocs.push_back(dn);
        if (dp) s->dev_allocs.push_back(dp);
        d.fast_nodes = dn;
        d.fast_prims = dp;
        return CR_OK;
    }
    FastTreeHost ft;
    build_fast_tree(s->elements, visible, ft);
    std::vector<uint2> table(ft.leaf_prims.size());
    for (size_t i = 0; i < table.size(); ++i) {
        const Element& e = s->elements[ft.leaf_prims[i]];
        table[i] = make_uint2(make_leaf(e.kind, e.idx), 0u);
    }
    void *dn = nullptr, *dp = nullptr, *dm = nullptr;
    int rc;
    if ((rc = upload(s, ft.nodes, &dn)) != CR_OK) return rc;
    if ((rc = upload(s, table, &dp)) != CR_OK) return rc;
    const uint32_t off_tri = d.n_prims[0], off_quad = d.n_prims[0] + d.n_prims[1];
    const size_t n_all = (size_t)d.n_prims[0] + d.n_prims[1] + d.n_prims[2];
    API_CUDA(cudaMallocAsync(&dm, std::max<size_t>(n_all, 1) * sizeof(uint32_t), s->stream));
    const int threads = 256;
    const int g1 = (int)std::min<size_t>(((size_t)d.n_nodes + threads - 1) / threads, (size_t)s->num_sms * 16);
    const int g2 = (int)std::min<size_t>((table.size() + threads - 1) / threads, (size_t)s->num_sms * 16);
    k_fast_leafmap<<<std::max(g1, 1), threads, 0, s->stream>>>(static_cast<const NodeRec<double>*>(d.nodes[0]), d.n_nodes,
                                                               static_cast<uint32_t*>(dm), off_tri, off_quad);
    k_fast_fill<<<std::max(g2, 1), threads, 0, s->stream>>>(static_cast<uint2*>(dp), (uint32_t)table.size(), static_cast<const uint32_t*>(dm), off_tri,
                                                            off_quad);
    API_CUDA(cudaGetLastError());
    API_CUDA(cudaFreeAsync(dm, s->stream));
    API_CUDA(cudaStreamSynchronize(s->stream));  // `ft` and `table` die with this scope
    d.fast_nodes = dn;
    d.fast_prims = dp;
    d.n_fast_nodes = (uint32_t)ft.nodes.size();
    d.n_fast_prims = (uint32_t)table.size();
    return CR_OK;
}

int upload_scene(CrScene* s) {
    API_CUDA(cudaSetDevice(s->device));
    s->free_device_scene();
    SceneDeviceData& d = s->dev;
    d.num_sms = s->num_sms;
    d.n_nodes = (uint32_t)s->n_nodes;
    d.n_prims[0] = (uint32_t)(s->spheres.size() / 4);
    d.n_prims[1] = (uint32_t)(s->tris.size() / 9);
    d.n_prims[2] = (uint32_t)(s->quads.size() / 9);
    d.sky_kind = s->sky_kind;
    d.sky_image = s->sky_image;
    d.clamp_colors = 1;
    d.strict_boxes = s->groups.empty() ? 0 : 1;
    d.max_radiance = 1.0;
    for (auto& m : s->mats)
        if (m.kind == CR_MAT_EMISSIVE) {
            d.clamp_colors = 0;
            for (int k = 0; k < 3; ++k)
                if (m.emit[k] > d.max_radiance) d.max_radiance = m.emit[k];
        }
    // nodes
    if (s->d_flat_nodes) {
        // device-built tree: the same flattening as below, by kernels (bvh_build.cu)
        void *p64 = nullptr, *p32 = nullptr;
        API_CUDA(cudaMallocAsync(&p64, s->n_nodes * sizeof(NodeRec<double>), s->stream));
        s->dev_allocs.push_back(p64);
        API_CUDA(cudaMallocAsync(&p32, s->n_nodes * sizeof(NodeRec<float>), s->stream));
        s->dev_allocs.push_back(p32);
        std::string err;
        const int rc = gpu_flatten_nodes(s->device, s->stream, s->d_flat_nodes, s->n_nodes, p64, p32, &d.bmax, &d.bsmall, err);
        if (rc != CR_OK) return fail(rc, err);
        d.nodes[0] = p64;
        d.nodes[1] = p32;
    } else {
        std::vector<NodeRec<double>> n64(s->nodes.size());
        std::vector<NodeRec<float>> n32(s->nodes.size());
        for (size_t i = 0; i < s->nodes.size(); ++i) {
            const FlatNode& n = s->nodes[i];
            NodeRec<double>& a = n64[i];
            a.xmin = n.box.lo[0]; a.xmax = n.box.hi[0];
            a.ymin = n.box.lo[1]; a.ymax = n.box.hi[1];
            a.zmin = n.box.lo[2]; a.zmax = n.box.hi[2];
            // device words: inner = (skip link, axis), leaf node = (left primitive, right primitive)
            const bool leafnode = ref_is_leaf(n.left);
            a.left = leafnode ? n.left : n.skip;
            a.right = leafnode ? n.right : (n.axis & 3u);
            a.pad0 = a.pad1 = 0;
            NodeRec<float>& b = n32[i];
            b.xmin = f32_down(n.box.lo[0]); b.xmax = f32_up(n.box.hi[0]);
            b.ymin = f32_down(n.box.lo[1]); b.ymax = f32_up(n.box.hi[1]);
            b.zmin = f32_down(n.box.lo[2]); b.zmax = f32_up(n.box.hi[2]);
            bool finite = true;
            for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(n.box.lo[k]) && std::isfinite(n.box.hi[k]);
            if ((n.axis & FLAT_ALWAYS_PASS) || !finite) {
                // no box test (member of a nested HitList), or the empty box of an empty nested list: the f32 copy is NaN, so
                // the conservative filter answers "undecided" and the walk looks at the node's flag / the exact f64 box
                if (n.axis & FLAT_ALWAYS_PASS) a.left |= ALWAYS_PASS_BIT;
                b.xmin = b.xmax = b.ymin = b.ymax = b.zmin = b.zmax = std::numeric_limits<float>::quiet_NaN();
            }
            b.left = a.left;
            b.right = a.right;
        }
        // per-node |coordinate| bound; the ~90 % smallest share the tight filter bound, the rest get BIGBOX_BIT
        std::vector<float> nb(s->nodes.size());
        for (size_t i = 0; i < s->nodes.size(); ++i) {
            double bm = 0.0;
            for (int k = 0; k < 3; ++k) bm = std::max(bm, std::max(std::fabs(s->nodes[i].box.lo[k]), std::fabs(s->nodes[i].box.hi[k])));
            nb[i] = std::isfinite(bm) ? f32_up(bm) : 0.f;  // an empty nested list has the empty box (+inf, -inf): never filtered
        }
        d.bmax = nb.empty() ? 0.f : *std::max_element(nb.begin(), nb.end());
        d.bsmall = d.bmax;
        if (!nb.empty()) {
            std::vector<float> sorted = nb;
            std::sort(sorted.begin(), sorted.end());
            d.bsmall = sorted[(size_t)((sorted.size() - 1) * 0.9)];
        }
        for (size_t i = 0; i < s->nodes.size(); ++i) {
            const uint32_t big = nb[i] > d.bsmall ? (1u << 27) : 0u;
            n64[i].left |= big;
            n32[i].left |= big;
        }
        int rc;
        if ((rc = upload(s, n64, &d.nodes[0])) != CR_OK) return rc;
        if ((rc = upload(s, n32, &d.nodes[1])) != CR_OK) return rc;
    }
    // spheres
    {
        const size_t n = s->spheres.size() / 4;
        std::vector<SphereRec<double>> a(n);
        std::vector<SphereRec<float>> b(n);
        for (size_t i = 0; i < n; ++i) {
            const double* p = &s->spheres[4 * i];
            a[i] = {p[0], p[1], p[2], p[3]};
            b[i] = {(float)p[0], (float)p[1], (float)p[2], (float)p[3]};
        }
        int rc;
        if ((rc = upload(s, a, &d.spheres[0])) != CR_OK) return rc;
        if ((rc = upload(s, b, &d.spheres[1])) != CR_OK) return rc;
    }
    // triangles: a, e1 = b - a, e2 = c - a (triangle.rs:99-100; x - y == x + (-y) bit for bit)
    if (s->commit_info.builder == CR_BVH_DEVICE && !s->tris.empty()) {
        const size_t n = s->tris.size() / 9;
        void *p64 = nullptr, *p32 = nullptr;
        API_CUDA(cudaMallocAsync(&p64, n * sizeof(TriRec<double>), s->stream));
        s->dev_allocs.push_back(p64);
        API_CUDA(cudaMallocAsync(&p32, n * sizeof(TriRec<float>), s->stream));
        s->dev_allocs.push_back(p32);
        std::string err;
        const int rc = gpu_flatten_tris(s->device, s->stream, s->tris.data(), n, p64, p32, err);
        if (rc != CR_OK) return fail(rc, err);
        d.tris[0] = p64;
        d.tris[1] = p32;
    } else {
        const size_t n = s->tris.size() / 9;
        std::vector<TriRec<double>> a(n);
        std::vector<TriRec<float>> b(n);
        for (size_t i = 0; i < n; ++i) {
            const double* p = &s->tris[9 * i];
            TriRec<double>& t = a[i];
            t.ax = p[0]; t.ay = p[1]; t.az = p[2];
            t.e1x = p[3] - p[0]; t.e1y = p[4] - p[1]; t.e1z = p[5] - p[2];
            t.e2x = p[6] - p[0]; t.e2y = p[7] - p[1]; t.e2z = p[8] - p[2];
            t.pad = 0.0;
            TriRec<float>& u = b[i];
            u.ax = (float)t.ax; u.ay = (float)t.ay; u.az = (float)t.az;
            u.e1x = (float)t.e1x; u.e1y = (float)t.e1y; u.e1z = (float)t.e1z;
            u.e2x = (float)t.e2x; u.e2y = (float)t.e2y; u.e2z = (float)t.e2z;
            u.pad0 = u.pad1 = u.pad2 = 0.f;
        }
        int rc;
        if ((rc = upload(s, a, &d.tris[0])) != CR_OK) return rc;
        if ((rc = upload(s, b, &d.tris[1])) != CR_OK) return rc;
    }
    // quads (EXTENSION): normal = unit(u x v), D = normal . Q, w = n / (n . n)
    {
        const size_t n = s->quads.size() / 9;
        std::vector<QuadRec<double>> a(n);
        std::vector<QuadRec<float>> b(n);
        for (size_t i = 0; i < n; ++i) {
            const double* p = &s->quads[9 * i];
            const double ux = p[3], uy = p[4], uz = p[5], vx = p[6], vy = p[7], vz = p[8];
            const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
            const double l2 = nx * nx + ny * ny + nz * nz;
            const double il = 1.0 / std::sqrt(l2);
            const double ux_ = il * nx, uy_ = il * ny, uz_ = il * nz;  // unit_vector: (1/len) * n
            const double dd = ux_ * p[0] + uy_ * p[1] + uz_ * p[2];
            const double nn = nx * nx + ny * ny + nz * nz;  // dot(n, n)
            const double inn = 1.0 / nn;
            QuadRec<double>& q = a[i];
            q.qx = p[0]; q.qy = p[1]; q.qz = p[2];
            q.ux = ux; q.uy = uy; q.uz = uz;
            q.vx = vx; q.vy = vy; q.vz = vz;
            q.nx = ux_; q.ny = uy_; q.nz = uz_;
            q.wx = inn * nx; q.wy = inn * ny; q.wz = inn * nz;
            q.d = dd;
            QuadRec<float>& f = b[i];
            f.qx = (float)q.qx; f.qy = (float)q.qy; f.qz = (float)q.qz;
            f.ux = (float)q.ux; f.uy = (float)q.uy; f.uz = (float)q.uz;
            f.vx = (float)q.vx; f.vy = (float)q.vy; f.vz = (float)q.vz;
            f.nx = (float)q.nx; f.ny = (float)q.ny; f.nz = (float)q.nz;
            f.wx = (float)q.wx; f.wy = (float)q.wy; f.wz = (float)q.wz;
            f.d = (float)q.d;
        }
        int rc;
        if ((rc = upload(s, a, &d.quads[0])) != CR_OK) return rc;
        if ((rc = upload(s, b, &d.quads[1])) != CR_OK) return rc;
    }
    // per-primitive metadata
    std::vector<int32_t> kind_of_mat(s->mats.size());
    for (size_t i = 0; i < s->mats.size(); ++i) {
        const CrMaterial& cm = s->mats[i];
        kind_of_mat[i] = cm.kind | ((cm.kind == CR_MAT_LAMBERTIAN && tex_needs_uv(s, cm.tex)) ? MATKIND_NEEDS_UV : 0);
    }
    for (int k = 0; k < 3; ++k) {
        HostVec<PrimMeta> m(s->mat_of[k].size());  // 160 MB for 10 M triangles: a cached block, filled on all host threads
        parallel_ranges(m.size(), [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i) {
                m[i].material = s->mat_of[k][i];
                m[i].prim_index = s->prim_of[k][i];
                m[i].obj_id = s->obj_of[k][i];
                m[i].mat_kind = kind_of_mat[(size_t)m[i].material];
            }
        });
        void* p = nullptr;
        int rc = upload(s, m, &p);
        if (rc != CR_OK) return rc;
        d.meta[k] = static_cast<PrimMeta*>(p);
    }
    // materials, textures
    {
        std::vector<DevMaterial> m(s->mats.size());
        for (size_t i = 0; i < m.size(); ++i) {
            const CrMaterial& c = s->mats[i];
            m[i].kind = c.kind;
            m[i].tex = c.tex;
            m[i].scatter_prob = c.scatter_prob;
            m[i].fuzz = c.fuzz;
            m[i].ior = c.ior;
            for (int k = 0; k < 3; ++k) {
                m[i].albedo[k] = c.albedo[k];
                m[i].emit[k] = c.emit[k];
            }
        }
        void* p = nullptr;
        int rc = upload(s, m, &p);
        if (rc != CR_OK) return rc;
        d.mats = static_cast<DevMaterial*>(p);
        std::vector<DevTexture> t(s->texs.size());
        for (size_t i = 0; i < t.size(); ++i) {
            const CrTexture& c = s->texs[i];
            t[i].kind = c.kind;
            t[i].even = c.even;
            t[i].odd = c.odd;
            t[i].image = c.image;
            t[i].inv_scale = c.inv_scale;
            for (int k = 0; k < 3; ++k) t[i].color[k] = c.color[k];
        }
        rc = upload(s, t, &p);
        if (rc != CR_OK) return rc;
        d.texs = static_cast<DevTexture*>(p);
    }
    // images -> uchar4 CUDA arrays + point-sampled texture objects
    {
        std::vector<DevImage> di(s->images.size());
        for (size_t i = 0; i < di.size(); ++i) {
            HostImage& im = s->images[i];
            static const bool prof = getenv("CRB_PROFILE_COMMIT") != nullptr;
            const auto tp0 = std::chrono::steady_clock::now();
            HostVec<uchar4> px((size_t)im.w * im.h);  // cached block, converted on all host threads (a sky image: 8 M texels)
            parallel_ranges(px.size(), [&](size_t a, size_t b) {
                for (size_t k = a; k < b; ++k) px[k] = make_uchar4(im.rgb[3 * k], im.rgb[3 * k + 1], im.rgb[3 * k + 2], 255);
            });
            const auto tp1 = std::chrono::steady_clock::now();
            const bool reused = image_cache(s->device).take(im.w, im.h, &im.arr, &im.tex);
            if (!reused) {
                cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
                API_CUDA(cudaMallocArray(&im.arr, &fmt, (size_t)im.w, (size_t)im.h));
            }
            const auto tp2 = std::chrono::steady_clock::now();
            API_CUDA(cudaMemcpy2DToArray(im.arr, 0, 0, px.data(), (size_t)im.w * sizeof(uchar4), (size_t)im.w * sizeof(uchar4),
                                         (size_t)im.h, cudaMemcpyHostToDevice));
            if (prof) {
                const auto tp3 = std::chrono::steady_clock::now();
                auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                fprintf(stderr, "crucible_b200 image %dx%d: convert %.2f ms, cudaMallocArray %.2f ms, copy %.2f ms\n", im.w, im.h, ms(tp0, tp1), ms(tp1, tp2), ms(tp2, tp3));
            }
            if (!im.tex) {
                cudaResourceDesc res;
                memset(&res, 0, sizeof(res));
                res.resType = cudaResourceTypeArray;
                res.res.array.array = im.arr;
                cudaTextureDesc td;
                memset(&td, 0, sizeof(td));
                td.addressMode[0] = cudaAddressModeClamp;
                td.addressMode[1] = cudaAddressModeClamp;
                td.filterMode = cudaFilterModePoint;
                td.readMode = cudaReadModeElementType;
                td.normalizedCoords = 0;
                API_CUDA(cudaCreateTextureObject(&im.tex, &res, &td, nullptr));
            }
            di[i].tex = im.tex;
            di[i].w = im.w;
            di[i].h = im.h;
        }
        void* p = nullptr;
        int rc = upload(s, di, &p);
        if (rc != CR_OK) return rc;
        d.images = static_cast<DevImage*>(p);
    }
    // object keyframes: one concatenated key table + a track (range) per sphere and per triangle vertex
    if (!s->anim.empty()) {
        std::vector<CrAnimKey> keys;
        std::vector<AnimTrack> st(s->spheres.size() / 4, AnimTrack{0u, 0u}), tt(s->tris.size() / 9 * 3, AnimTrack{0u, 0u});
        std::vector<uint32_t> slot(s->tris.size() / 9, 0u);
        std::vector<double> verts;
        std::vector<char> tri_seen(s->tris.size() / 9, 0);
        for (auto& kv : s->anim) {
            if (kv.second.empty()) continue;
            const size_t prim = (size_t)(kv.first >> 2);
            const int point = (int)(kv.first & 3u);
            const Element& e = s->elements[prim];
            const AnimTrack tr{(uint32_t)keys.size(), (uint32_t)kv.second.size()};
            keys.insert(keys.end(), kv.second.begin(), kv.second.end());
            if (e.kind == CR_PRIM_SPHERE) {
                st[e.idx] = tr;
            } else {
                tt[3 * (size_t)e.idx + (size_t)point] = tr;
                if (!tri_seen[e.idx]) {
                    tri_seen[e.idx] = 1;
                    slot[e.idx] = (uint32_t)(verts.size() / 9);
                    verts.insert(verts.end(), &s->tris[9 * (size_t)e.idx], &s->tris[9 * (size_t)e.idx] + 9);
                }
            }
        }
        if (!keys.empty()) {
            void* p = nullptr;
            int rc;
            if ((rc = upload(s, keys, &p)) != CR_OK) return rc;
            d.anim_keys = static_cast<CrAnimKey*>(p);
            if ((rc = upload(s, st, &p)) != CR_OK) return rc;
            d.sphere_track = static_cast<AnimTrack*>(p);
            if ((rc = upload(s, tt, &p)) != CR_OK) return rc;
            d.tri_track = static_cast<AnimTrack*>(p);
            if ((rc = upload(s, slot, &p)) != CR_OK) return rc;
            d.tri_anim_slot = static_cast<uint32_t*>(p);
            if ((rc = upload(s, verts, &p)) != CR_OK) return rc;
            d.tri_anim_verts = static_cast<double*>(p);
        }
    }
    // the render may run on a caller-supplied stream: the scene is complete when commit returns
    API_CUDA(cudaStreamSynchronize(s->stream));
    return CR_OK;
}

int validate(CrScene* s) {
    const int nm = (int)s->mats.size(), nt = (int)s->texs.size(), ni = (int)s->images.size();
    // the shading queues carry the material index in 24 bits (bit 31 = "needs u, v"), integrator.cuh RenderTraceIO::commit
    if (s->mats.size() > (1u << 24)) return fail(CR_ERR_LIMIT, "more than 16 777 216 materials");
    for (int k = 0; k < 3; ++k)
        for (int32_t m : s->mat_of[k])
            if (m < 0 || m >= nm) return fail(CR_ERR_INVALID, "primitive references material " + std::to_string(m) + " of " + std::to_string(nm));
    for (auto& m : s->mats) {
        if (m.kind < CR_MAT_LAMBERTIAN || m.kind > CR_MAT_EMISSIVE) return fail(CR_ERR_INVALID, "bad material kind");
        if (m.kind == CR_MAT_LAMBERTIAN) {
            if (m.tex < 0 || m.tex >= nt) return fail(CR_ERR_INVALID, "lambertian references a missing texture");
            if (!(m.scatter_prob > 0.0)) return fail(CR_ERR_INVALID, "lambertian scatter_prob must be > 0");
        }
        // Metal::new asserts fuzz in [0,1] (metal.rs:21-25); Color::new asserts [0,1] (utils.rs:345-351)
        if (m.kind == CR_MAT_METAL) {
            if (!(m.fuzz >= 0.0 && m.fuzz <= 1.0)) return fail(CR_ERR_INVALID, "A metal cannot have a fuzz factor outside [0,1]");
            for (int k = 0; k < 3; ++k)
                if (!(m.albedo[k] >= 0.0 && m.albedo[k] <= 1.0)) return fail(CR_ERR_INVALID, "metal albedo must be in [0,1]");
        }
    }
    for (auto& t : s->texs) {
        if (t.kind == CR_TEX_SOLID) {
            for (int k = 0; k < 3; ++k)
                if (!(t.color[k] >= 0.0 && t.color[k] <= 1.0)) return fail(CR_ERR_INVALID, "solid colour must be in [0,1]");
        } else if (t.kind == CR_TEX_CHECKER) {
            if (t.even < 0 || t.even >= nt || t.odd < 0 || t.odd >= nt) return fail(CR_ERR_INVALID, "checker references a missing texture");
        } else if (t.kind == CR_TEX_IMAGE) {
            if (t.image < 0 || t.image >= ni) return fail(CR_ERR_INVALID, "image texture references a missing image");
        } else {
            return fail(CR_ERR_INVALID, "bad texture kind");
        }
    }
    // checker nesting must terminate within MAX_TEX_NEST
    for (int i = 0; i < nt; ++i) {
        std::vector<int> frontier = {i};
        for (int depth = 0; depth <= MAX_TEX_NEST && !frontier.empty(); ++depth) {
            std::vector<int> next;
            for (int t : frontier)
                if (s->texs[(size_t)t].kind == CR_TEX_CHECKER) {
                    next.push_back(s->texs[(size_t)t].even);
                    next.push_back(s->texs[(size_t)t].odd);
                }
            if (depth == MAX_TEX_NEST - 1 && !next.empty()) return fail(CR_ERR_LIMIT, "checker textures nested deeper than 8");
            frontier.swap(next);
        }
    }
    if (s->sky_kind == CR_SKY_SPHERICAL && (s->sky_image < 0 || s->sky_image >= ni))
        return fail(CR_ERR_INVALID, "spherical sky references a missing image");
    return CR_OK;
}

int need_device(CrScene* s) {
    if (!s) return fail(CR_ERR_INVALID, "null scene");
    if (s->device < 0) return fail(CR_ERR_NO_DEVICE, "no CUDA sm_100 device: crucible_b200 has no CPU fallback");
    if (!s->committed) return fail(CR_ERR_STATE, "scene not committed (call cr_scene_commit)");
    return CR_OK;
}

}  // namespace

extern "C" {

const char* cr_last_error(void) { return g_err.c_str(); }
const char* cr_version(void) { return "crucible_b200 0.1 (sm_100a)"; }

int cr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n && i < 64; ++i)
        if (device_info(i).state == 1) ++ok;
    return ok;
}

CrScene* cr_scene_create(int device) {
    CrScene* s = new CrScene();
    if (device >= 0) {
        const DeviceInfo& di = device_info(device);
        if (di.state != 1) {
            g_err = "cr_scene_create: " + di.why;
            delete s;
            return nullptr;
        }
        s->device = device;
        s->num_sms = di.num_sms;
        cudaSetDevice(device);
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
            g_err = "cr_scene_create: cudaStreamCreate failed";
            delete s;
            return nullptr;
        }
    }
    return s;
}

int cr_device_trim(int device) {
    if (device < 0) {  // host-only scenes (cr_scene_create(-1)): only the staging cache exists
        HostBlockPool::instance().trim();
        return CR_OK;
    }
    const DeviceInfo& di = device_info(device);
    if (di.state != 1) return fail(CR_ERR_NO_DEVICE, di.why);
    DeviceSlot& slot = device_slot(device);
    std::lock_guard<std::mutex> lk(slot.mu);
    API_CUDA(cudaSetDevice(device));
    API_CUDA(cudaDeviceSynchronize());
    slot.ws.release();
    image_cache(device).trim();
    HostBlockPool::instance().trim();  // host staging blocks cached between scenes (host_pool.h)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = 0;  // the CUDA default: cached blocks go back to the driver at the next synchronisation
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        cudaMemPoolTrimTo(pool, 0);
        keep = UINT64_MAX;  // later scenes of this process cache again
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    return CR_OK;
}

void cr_scene_destroy(CrScene* s) {
    if (!s) return;
    if (s->device >= 0) {
        static const bool prof = getenv("CRB_PROFILE_COMMIT") != nullptr;
        auto ms = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
        const auto t0 = std::chrono::steady_clock::now();
        cudaSetDevice(s->device);
        s->free_device_scene();
        s->free_flat_tree();
        const double ms_scene = ms(t0);
        if (s->d_out_rgb) cudaFreeAsync(s->d_out_rgb, s->stream);
        if (s->d_out_rgb8) cudaFreeAsync(s->d_out_rgb8, s->stream);
        if (s->d_io) cudaFreeAsync(s->d_io, s->stream);
        if (s->d_multi_rgb) cudaFree(s->d_multi_rgb);
        if (s->d_multi_rgb8) cudaFree(s->d_multi_rgb8);
        const double ms_bufs = ms(t0);
        if (s->stream) cudaStreamDestroy(s->stream);  // resources are released once the queued work has drained
        if (prof) fprintf(stderr, "crucible_b200 destroy: device scene %.2f ms, buffers %.2f ms, stream %.2f ms\n", ms_scene, ms_bufs - ms_scene, ms(t0) - ms_bufs);
    }
    const auto t1 = std::chrono::steady_clock::now();
    delete s;
    if (getenv("CRB_PROFILE_COMMIT")) fprintf(stderr, "crucible_b200 destroy: host %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
}

// Box of one primitive in reference arithmetic.
static inline Box prim_box(uint32_t kind, const double* p) {
    Box box;
    if (kind == CR_PRIM_SPHERE) {
        // sphere.rs:29-30: new_from_points(center - rvec, center + rvec); a - b == a + (-b)
        const double lo[3] = {p[0] - p[3], p[1] - p[3], p[2] - p[3]};
        const double hi[3] = {p[0] + p[3], p[1] + p[3], p[2] + p[3]};
        box = box_from_points(lo, hi);
    } else if (kind == CR_PRIM_TRIANGLE) {
        // triangle.rs:48-62: a.min(b.min(c)) / a.max(b.max(c)) per axis (f64::min/max ignore NaN)
        for (int k = 0; k < 3; ++k) {
            box.lo[k] = std::fmin(p[k], std::fmin(p[3 + k], p[6 + k]));
            box.hi[k] = std::fmax(p[k], std::fmax(p[3 + k], p[6 + k]));
        }
    } else {
        // EXTENSION quad: box of both diagonals, sides thinner than 1e-4 padded (Interval::pad, utils.rs:624-627)
        double quv[3], qu[3], qv[3];
        for (int k = 0; k < 3; ++k) {
            qu[k] = p[k] + p[3 + k];
            qv[k] = p[k] + p[6 + k];
            quv[k] = qu[k] + p[6 + k];
        }
        box = box_union(box_from_points(p, quv), box_from_points(qu, qv));
        const double delta = 0.0001;
        for (int k = 0; k < 3; ++k)
            if (box.hi[k] - box.lo[k] < delta) {
                const double pad = delta / 2.0;
                box.lo[k] = box.lo[k] - pad;
                box.hi[k] = box.hi[k] + pad;
            }
    }
    return box;
}

// Appends n primitives of one kind (Scene::add_element / load_asset order).  The batch is validated first, so a
// rejected batch leaves the scene untouched; the staging arrays grow once and big batches (meshes of millions
// of triangles) are boxed by a few host threads.
static int64_t add_prims(CrScene* s, uint32_t kind, const double* data, size_t stride, const int32_t* material,
                         const int32_t* obj_id, size_t n) {
    if (!s || (!data && n)) return fail(CR_ERR_INVALID, "null argument");
    const size_t first = s->elements.size();
    if (first + n > REF_MAX_INDEX) return fail(CR_ERR_LIMIT, "too many primitives");
    {  // all coordinates finite: exponent field != 0x7ff, as a branch-free integer reduction (vectorises; a mesh world arrives
       // as 1.4 GB through this loop)
        uint64_t bad = 0;
        for (size_t k = 0; k < n * stride; ++k) {
            uint64_t bits;
            std::memcpy(&bits, &data[k], sizeof bits);
            bad |= (uint64_t)(((bits >> 52) & 0x7ffu) == 0x7ffu);
        }
        if (bad) return fail(CR_ERR_INVALID, "non-finite primitive coordinate");
    }
    if (kind == CR_PRIM_SPHERE)
        for (size_t i = 0; i < n; ++i)
            if (!(data[4 * i + 3] >= 0.0)) return fail(CR_ERR_INVALID, "Cannot make a sphere with negative radius");  // sphere.rs:26
    HostVec<double>& store = kind == CR_PRIM_SPHERE ? s->spheres : (kind == CR_PRIM_TRIANGLE ? s->tris : s->quads);
    const size_t first_of_kind = store.size() / stride;
    HostVec<int32_t>&mo = s->mat_of[kind], &oo = s->obj_of[kind], &po = s->prim_of[kind];
    // The boxes (Sphere::new / Triangle::new, sphere.rs:29-30, triangle.rs:27-35) are computed at commit, for every new
    // element at once and on all host threads: a mesh scene arrives as thousands of add calls (Scene::load_asset adds one
    // mesh per call), and boxing them call by call on one thread cost 1.6 s of the 10 M-triangle scene's 2 s staging.
    // Every staging array is written ONCE here (no resize-then-fill: 1.5 GB pass through this function for that scene).
    try {
        store.insert(store.end(), data, data + n * stride);
        for (size_t i = 0; i < n; ++i) {
            Element e;
            e.kind = kind;
            e.idx = (uint32_t)(first_of_kind + i);
            e.hide = false;
            s->elements.push_back(e);
        }
        if (material) mo.insert(mo.end(), material, material + n);
        else mo.resize(first_of_kind + n, 0);
        if (obj_id) oo.insert(oo.end(), obj_id, obj_id + n);
        else
            for (size_t i = 0; i < n; ++i) oo.push_back((int32_t)(first + i));
        for (size_t i = 0; i < n; ++i) po.push_back((int32_t)(first + i));
    } catch (const std::bad_alloc&) {  // leave the scene as it was
        store.resize(first_of_kind * stride);
        s->elements.resize(first);
        mo.resize(first_of_kind);
        oo.resize(first_of_kind);
        po.resize(first_of_kind);
        return fail(CR_ERR_INVALID, "out of host memory while staging primitives");
    }
    if (!s->groups.empty()) {  // member lists exist once the scene has a group
        std::vector<uint32_t>& dst = s->open_groups.empty() ? s->top : s->groups[(size_t)s->open_groups.back()].members;
        for (size_t i = 0; i < n; ++i) dst.push_back((uint32_t)(first + i));
    }
    s->committed = false;
    return (int64_t)first;
}

// Many add calls in one: batch b appends counts[b] primitives of kind kinds[b], exactly as the matching cr_scene_add_* call
// would, in order.  A mesh world is one batch per mesh (Scene::load_asset, scene/mod.rs:211-229: ~1 600 for the 10 M-triangle
// scene); handing them over together lets the library validate and copy them on all host threads into arrays sized once
// (0.39 s -> 0.1 s for that scene's 1.5 GB).  All batches are validated before anything is appended.
int64_t cr_scene_add_batches(CrScene* s, size_t n_batches, const int32_t* kinds, const double* const* data, const int32_t* const* material,
                             const int32_t* const* obj_id, const size_t* counts) {
    if (!s || (n_batches && (!kinds || !data || !counts))) return fail(CR_ERR_INVALID, "null argument");
    const size_t first = s->elements.size();
    // offsets of every batch: in the flat element list and in its kind's arrays
    std::vector<size_t> at(n_batches), at_kind(n_batches);
    size_t n_kind[3] = {s->spheres.size() / 4, s->tris.size() / 9, s->quads.size() / 9};
    const size_t old_kind[3] = {n_kind[0], n_kind[1], n_kind[2]};
    size_t total = first;
    for (size_t b = 0; b < n_batches; ++b) {
        if (kinds[b] < CR_PRIM_SPHERE || kinds[b] > CR_PRIM_QUAD) return fail(CR_ERR_INVALID, "bad primitive kind");
        if (!data[b] && counts[b]) return fail(CR_ERR_INVALID, "null argument");
        at[b] = total;
        at_kind[b] = n_kind[kinds[b]];
        total += counts[b];
        n_kind[kinds[b]] += counts[b];
        if (total > REF_MAX_INDEX) return fail(CR_ERR_LIMIT, "too many primitives");
    }
    // validation, batches in parallel: 0 = fine, 1 = non-finite coordinate, 2 = negative radius
    std::atomic<int> bad{0};
    parallel_ranges_always(n_batches, total - first, [&](size_t b0, size_t b1) {
        for (size_t b = b0; b < b1 && bad.load(std::memory_order_relaxed) == 0; ++b) {
            const size_t stride = kinds[b] == CR_PRIM_SPHERE ? 4 : 9, n = counts[b];
            const double* d = data[b];
            uint64_t nonfinite = 0;
            for (size_t k = 0; k < n * stride; ++k) {
                uint64_t bits;
                std::memcpy(&bits, &d[k], sizeof bits);
                nonfinite |= (uint64_t)(((bits >> 52) & 0x7ffu) == 0x7ffu);
            }
            if (nonfinite) bad.store(1);
            if (kinds[b] == CR_PRIM_SPHERE)
                for (size_t i = 0; i < n; ++i)
                    if (!(d[4 * i + 3] >= 0.0)) bad.store(2);  // sphere.rs:26
        }
    });
    if (bad.load() == 1) return fail(CR_ERR_INVALID, "non-finite primitive coordinate");
    if (bad.load() == 2) return fail(CR_ERR_INVALID, "Cannot make a sphere with negative radius");
    if (!s->groups.empty()) {  // scenes with nested elements keep member lists: the plain calls maintain them (all batches are valid)
        for (size_t b = 0; b < n_batches; ++b) {
            const int64_t rc = add_prims(s, (uint32_t)kinds[b], data[b], kinds[b] == CR_PRIM_SPHERE ? 4 : 9, material ? material[b] : nullptr,
                                         obj_id ? obj_id[b] : nullptr, counts[b]);
            if (rc < 0) return rc;  // out of memory only
        }
        return (int64_t)first;
    }
    HostVec<double>* stores[3] = {&s->spheres, &s->tris, &s->quads};
    try {  // sized once; HostVec leaves the new tail uninitialised (host_pool.h), every byte is written below
        s->elements.resize(total);
        for (int k = 0; k < 3; ++k) {
            stores[k]->resize(n_kind[k] * (k == 0 ? 4 : 9));
            s->mat_of[k].resize(n_kind[k]);
            s->obj_of[k].resize(n_kind[k]);
            s->prim_of[k].resize(n_kind[k]);
        }
    } catch (const std::bad_alloc&) {
        s->elements.resize(first);
        for (int k = 0; k < 3; ++k) {
            stores[k]->resize(old_kind[k] * (k == 0 ? 4 : 9));
            s->mat_of[k].resize(old_kind[k]);
            s->obj_of[k].resize(old_kind[k]);
            s->prim_of[k].resize(old_kind[k]);
        }
        return fail(CR_ERR_INVALID, "out of host memory while staging primitives");
    }
    parallel_ranges_always(n_batches, total - first, [&](size_t b0, size_t b1) {
        for (size_t b = b0; b < b1; ++b) {
            const int k = kinds[b];
            const size_t stride = k == CR_PRIM_SPHERE ? 4 : 9, n = counts[b], e0 = at[b], k0 = at_kind[b];
            if (n) std::memcpy(stores[k]->data() + k0 * stride, data[b], n * stride * sizeof(double));
            const int32_t* m = material ? material[b] : nullptr;
            const int32_t* o = obj_id ? obj_id[b] : nullptr;
            Element* el = s->elements.data() + e0;
            int32_t *mo = s->mat_of[k].data() + k0, *oo = s->obj_of[k].data() + k0, *po = s->prim_of[k].data() + k0;
            for (size_t i = 0; i < n; ++i) {
                el[i].kind = (uint32_t)k;
                el[i].idx = (uint32_t)(k0 + i);
                el[i].hide = false;
                mo[i] = m ? m[i] : 0;
                oo[i] = o ? o[i] : (int32_t)(e0 + i);
                po[i] = (int32_t)(e0 + i);
            }
        }
    });
    s->committed = false;
    return (int64_t)first;
}

int cr_scene_reserve(CrScene* s, size_t n_spheres, size_t n_triangles, size_t n_quads) {
    if (!s) return fail(CR_ERR_INVALID, "null scene");
    const size_t add[3] = {n_spheres, n_triangles, n_quads};
    const size_t total = n_spheres + n_triangles + n_quads;
    if (s->elements.size() + total > REF_MAX_INDEX) return fail(CR_ERR_LIMIT, "too many primitives");
    try {
        s->elements.reserve(s->elements.size() + total);
        s->spheres.reserve(s->spheres.size() + 4 * n_spheres);
        s->tris.reserve(s->tris.size() + 9 * n_triangles);
        s->quads.reserve(s->quads.size() + 9 * n_quads);
        for (int k = 0; k < 3; ++k) {
            s->mat_of[k].reserve(s->mat_of[k].size() + add[k]);
            s->obj_of[k].reserve(s->obj_of[k].size() + add[k]);
            s->prim_of[k].reserve(s->prim_of[k].size() + add[k]);
        }
    } catch (const std::bad_alloc&) {
        return fail(CR_ERR_INVALID, "cr_scene_reserve: out of host memory");
    }
    return CR_OK;
}

int cr_scene_begin_group(CrScene* s, int kind) {
    if (!s || (kind != CR_GROUP_HITLIST && kind != CR_GROUP_BVH)) return fail(CR_ERR_INVALID, "bad group kind");
    if (s->groups.size() >= 0x7FFFFFFEu) return fail(CR_ERR_LIMIT, "too many groups");
    if (s->groups.empty()) {  // everything added so far is a top-level element
        s->top.resize(s->elements.size());
        for (size_t i = 0; i < s->top.size(); ++i) s->top[i] = (uint32_t)i;
    }
    const int32_t id = (int32_t)s->groups.size();
    const int32_t parent = s->open_groups.empty() ? -1 : s->open_groups.back();
    (parent < 0 ? s->top : s->groups[(size_t)parent].members).push_back(GROUP_MEMBER | (uint32_t)id);
    s->groups.push_back(CrScene::Group{kind, parent, {}});
    s->open_groups.push_back(id);
    s->committed = false;
    return id;
}
int cr_scene_end_group(CrScene* s) {
    if (!s || s->open_groups.empty()) return fail(CR_ERR_STATE, "cr_scene_end_group without an open group");
    s->open_groups.pop_back();
    return CR_OK;
}

int64_t cr_scene_add_spheres(CrScene* s, const double* d, const int32_t* m, const int32_t* o, size_t n) {
    return add_prims(s, CR_PRIM_SPHERE, d, 4, m, o, n);
}
int64_t cr_scene_add_triangles(CrScene* s, const double* d, const int32_t* m, const int32_t* o, size_t n) {
    return add_prims(s, CR_PRIM_TRIANGLE, d, 9, m, o, n);
}
int64_t cr_scene_add_quads(CrScene* s, const double* d, const int32_t* m, const int32_t* o, size_t n) {
    return add_prims(s, CR_PRIM_QUAD, d, 9, m, o, n);
}

int cr_scene_set_hidden(CrScene* s, size_t prim_index, int hide) {
    if (!s || prim_index >= s->elements.size()) return fail(CR_ERR_INVALID, "prim_index out of range");
    s->elements[prim_index].hide = hide != 0;
    s->committed = false;
    return CR_OK;
}
int cr_scene_set_materials(CrScene* s, const CrMaterial* m, size_t n) {
    if (!s || (!m && n)) return fail(CR_ERR_INVALID, "null argument");
    s->mats.assign(m, m + n);
    s->committed = false;
    return CR_OK;
}
int cr_scene_set_textures(CrScene* s, const CrTexture* t, size_t n) {
    if (!s || (!t && n)) return fail(CR_ERR_INVALID, "null argument");
    s->texs.assign(t, t + n);
    s->committed = false;
    return CR_OK;
}
int cr_scene_add_image(CrScene* s, const uint8_t* rgb8, int w, int h) {
    if (!s || !rgb8 || w <= 0 || h <= 0) return fail(CR_ERR_INVALID, "bad image");
    HostImage im;
    im.w = w;
    im.h = h;
    im.rgb.assign(rgb8, rgb8 + (size_t)w * h * 3);
    s->images.push_back(std::move(im));
    s->committed = false;
    return (int)s->images.size() - 1;
}
int cr_scene_set_sky(CrScene* s, int kind, int image) {
    if (!s || kind < CR_SKY_DEFAULT || kind > CR_SKY_BLACK) return fail(CR_ERR_INVALID, "bad sky kind");
    s->sky_kind = kind;
    s->sky_image = image;
    s->committed = false;
    return CR_OK;
}

int cr_scene_set_keyframes(CrScene* s, size_t prim_index, int point, const CrAnimKey* keys, size_t n) {
    if (!s || (!keys && n)) return fail(CR_ERR_INVALID, "null argument");
    if (prim_index >= s->elements.size()) return fail(CR_ERR_INVALID, "prim_index out of range");
    const uint32_t kind = s->elements[prim_index].kind;
    if (kind == CR_PRIM_QUAD) return fail(CR_ERR_INVALID, "quads (extension) cannot be animated");
    if (point < 0 || point > (kind == CR_PRIM_SPHERE ? 0 : 2)) return fail(CR_ERR_INVALID, "point out of range for this primitive");
    for (size_t i = 0; i < n; ++i) {
        if (keys[i].kind < 0 || keys[i].kind > 6) return fail(CR_ERR_INVALID, "keyframe kind out of range");
        // ScaleR can only be applied to Spheres (scene_animator.rs:140-150)
        if (keys[i].kind == 3 && kind != CR_PRIM_SPHERE) return fail(CR_ERR_INVALID, "ScaleR can only be applied to Spheres");
        // scale_x / scale_y / scale_z reject spheres (scene_animator.rs:38-41, 72-75, 106-109)
        if (keys[i].kind >= 4 && kind == CR_PRIM_SPHERE)
            return fail(CR_ERR_INVALID, keys[i].kind == 4   ? "ScaleX cannot apply to Spheres"
                                        : keys[i].kind == 5 ? "ScaleY cannot apply to Spheres"
                                                            : "ScaleZ cannot apply to Spheres");
        if (keys[i].interp != CR_NERP && keys[i].interp != CR_LERP) return fail(CR_ERR_INVALID, "bad interpolation type");
    }
    std::vector<CrAnimKey>& dst = s->anim[((uint64_t)prim_index << 2) | (uint64_t)point];
    dst.assign(keys, keys + n);
    s->committed = false;
    return CR_OK;
}

// Boxes of the elements added since the last commit, in parallel.
static void ensure_boxes(CrScene* s) {
    const size_t a0 = s->boxed, a1 = s->elements.size();
    if (a0 >= a1) return;
    auto fill = [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            Element& e = s->elements[i];
            const HostVec<double>& store = e.kind == CR_PRIM_SPHERE ? s->spheres : (e.kind == CR_PRIM_TRIANGLE ? s->tris : s->quads);
            e.box = prim_box(e.kind, &store[(size_t)(e.kind == CR_PRIM_SPHERE ? 4 : 9) * e.idx]);
        }
    };
    parallel_ranges(a1 - a0, [&](size_t a, size_t b) { fill(a0 + a, a0 + b); });
    s->boxed = a1;
}

int cr_scene_commit(CrScene* s) {
    if (!s) return fail(CR_ERR_INVALID, "null scene");
    int rc = validate(s);
    if (rc != CR_OK) return rc;
    ensure_boxes(s);
    if (!s->open_groups.empty()) return fail(CR_ERR_STATE, "a group is still open (cr_scene_end_group missing)");
    const bool grouped = !s->groups.empty();
    // BVHWrapper::new_wrapper: drop hidden primitives, empty -> empty HitList (bvhwrapper.rs:16-31)
    std::vector<uint32_t> visible;
    if (!grouped) {
        visible.reserve(s->elements.size());
        for (uint32_t i = 0; i < (uint32_t)s->elements.size(); ++i)
            if (!s->elements[i].hide) visible.push_back(i);
    }
    s->n_visible = visible.size();
    s->nodes.clear();
    s->n_nodes = 0;
    if (s->device >= 0) {
        API_CUDA(cudaSetDevice(s->device));
        s->free_flat_tree();
    }
    s->root = REF_MISS;
    s->max_depth = 0;
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };
    const auto t_commit = clk::now();
    s->commit_info = CrCommitInfo{};
    s->commit_info.builder = CR_BVH_HOST;
    if (s->bvh_builder == CR_BVH_DEVICE && s->device < 0) return fail(CR_ERR_NO_DEVICE, "CR_BVH_DEVICE needs a scene created on a CUDA device");
    // DEVICE: tree build, record flattening and triangle set-up run as kernels; HOST: on the host, as written
    const bool on_device = !grouped && (s->bvh_builder == CR_BVH_DEVICE || (s->bvh_builder == CR_BVH_AUTO && s->device >= 0 && visible.size() >= 32768));
    if (on_device) s->commit_info.builder = CR_BVH_DEVICE;
    s->leaf_order_grouped.clear();
    if (grouped) {
        // nested elements: one preorder array for the whole Hittables tree (GroupEmitter), built on the host
        GroupEmitter em{*s, s->nodes};
        em.compute_boxes();
        std::vector<uint32_t> vis = em.visible_of(s->top);
        if (!vis.empty()) em.emit_bvh(vis, 0, vis.size(), 1);
        if (s->nodes.size() > REF_MAX_INDEX) return fail(CR_ERR_LIMIT, "BVH too large");
        s->n_nodes = s->nodes.size();
        s->max_depth = em.max_depth;
        s->n_visible = em.leaf_order.size();
        s->leaf_order_grouped.swap(em.leaf_order);
        if (s->n_nodes) s->root = 0;
        s->commit_info.levels = s->max_depth;
        s->commit_info.ms_build = ms_since(t_commit);
    } else if (!visible.empty()) {
        const uint64_t nn = node_count(visible.size());
        if (nn > REF_MAX_INDEX) return fail(CR_ERR_LIMIT, "BVH too large");
        if (on_device) {
            BvhBuildTimes bt;
            std::string err;
            rc = gpu_build_bvh(s->device, s->stream, s->elements, visible, &s->d_flat_nodes, &s->n_nodes, s->max_depth, &bt, err);
            if (rc != CR_OK) return fail(rc, err);
            s->commit_info.ms_pack = bt.ms_pack;
            s->commit_info.ms_h2d = bt.ms_h2d;
            s->commit_info.ms_device = bt.ms_device;
        } else {
            s->nodes.resize((size_t)nn);
            s->n_nodes = nn;
            Builder b{*s, visible, s->nodes};
            s->max_depth = b.build(0, visible.size(), 0, 4);
            // new_from_vec (bvhwrapper.rs:34-44): the root box is re-derived from its two children
            FlatNode& r = s->nodes[0];
            auto child_box = [&](uint32_t ref) -> Box {
                if (ref_is_leaf(ref)) return s->elements[(size_t)s->prim_of[ref_kind(ref)][ref_index(ref)]].box;
                return s->nodes[ref].box;
            };
            const Box lb = child_box(r.left);
            const Box rb = (r.right == REF_NONE) ? lb : child_box(r.right);
            r.box = box_union(lb, rb);
        }
        s->root = 0;
        s->commit_info.levels = s->max_depth;
        s->commit_info.ms_build = ms_since(t_commit);
    }
    s->committed = true;
    if (s->device >= 0) {
        const auto t_upload = clk::now();
        rc = upload_scene(s);
        if (rc != CR_OK) {
            s->committed = false;
            return rc;
        }
        s->commit_info.ms_upload = ms_since(t_upload);
        const auto t_search = clk::now();
        rc = upload_search_tree(s, visible);
        if (rc != CR_OK) {
            s->committed = false;
            return rc;
        }
        s->commit_info.ms_search_tree = ms_since(t_search);
    }
    s->commit_info.ms_total = ms_since(t_commit);
    return CR_OK;
}

int cr_scene_set_bvh_builder(CrScene* s, int builder) {
    if (!s || builder < CR_BVH_AUTO || builder > CR_BVH_DEVICE) return fail(CR_ERR_INVALID, "bad BVH builder");
    s->bvh_builder = builder;
    s->committed = false;
    return CR_OK;
}

int cr_scene_commit_info(const CrScene* s, CrCommitInfo* out) {
    if (!s || !out) return fail(CR_ERR_INVALID, "null argument");
    if (!s->committed) return fail(CR_ERR_STATE, "scene not committed");
    *out = s->commit_info;
    return CR_OK;
}

// host copy of the tree; a device-built tree is fetched the first time somebody looks at it
static int host_nodes(const CrScene* cs) {
    CrScene* s = const_cast<CrScene*>(cs);
    if (!s->d_flat_nodes || s->nodes.size() == s->n_nodes) return CR_OK;
    std::string err;
    const int rc = gpu_fetch_flat_nodes(s->device, s->stream, s->d_flat_nodes, s->n_nodes, s->nodes, err);
    return rc == CR_OK ? rc : fail(rc, err);
}

int64_t cr_scene_bvh_nodes(const CrScene* s, CrBvhNode* out, size_t cap) {
    if (!s || !s->committed) return fail(CR_ERR_STATE, "scene not committed");
    if (!out || !cap) return (int64_t)s->n_nodes;
    const int rc = host_nodes(s);
    if (rc != CR_OK) return rc;
    const size_t n = std::min(cap, s->nodes.size());
    for (size_t i = 0; out && i < n; ++i) {
        const FlatNode& f = s->nodes[i];
        CrBvhNode& o = out[i];
        for (int k = 0; k < 3; ++k) {
            o.lo[k] = f.box.lo[k];
            o.hi[k] = f.box.hi[k];
        }
        o.left = f.left;
        o.right = f.right;
        o.axis = f.axis;
        o.skip = f.skip;
    }
    return (int64_t)s->nodes.size();
}

int64_t cr_scene_device_records(const CrScene* s, int which, void* out, size_t cap_bytes) {
    if (!s || !s->committed) return fail(CR_ERR_STATE, "scene not committed");
    if (s->device < 0) return fail(CR_ERR_NO_DEVICE, "the scene has no device copy");
    if (which < 0 || which > 3) return fail(CR_ERR_INVALID, "which: 0 nodes f64, 1 nodes f32, 2 triangles f64, 3 triangles f32");
    const size_t n_tris = s->tris.size() / 9;
    const size_t bytes = which == 0   ? s->n_nodes * sizeof(NodeRec<double>)
                         : which == 1 ? s->n_nodes * sizeof(NodeRec<float>)
                         : which == 2 ? n_tris * sizeof(TriRec<double>)
                                      : n_tris * sizeof(TriRec<float>);
    const void* src = which < 2 ? s->dev.nodes[which] : s->dev.tris[which - 2];
    const size_t n = std::min(bytes, cap_bytes);
    if (out && n && src) {
        API_CUDA(cudaSetDevice(s->device));
        API_CUDA(cudaMemcpyAsync(out, src, n, cudaMemcpyDeviceToHost, s->stream));
        API_CUDA(cudaStreamSynchronize(s->stream));
    }
    return (int64_t)bytes;
}

int cr_scene_bvh_info(const CrScene* s, uint64_t* n_nodes, uint32_t* max_depth, uint64_t* n_visible) {
    if (!s || !s->committed) return fail(CR_ERR_STATE, "scene not committed");
    if (n_nodes) *n_nodes = s->n_nodes;
    if (max_depth) *max_depth = s->max_depth;
    if (n_visible) *n_visible = s->n_visible;
    return CR_OK;
}

int64_t cr_scene_bvh_leaf_order(const CrScene* s, int32_t* out, size_t cap) {
    if (!s || !s->committed) return fail(CR_ERR_STATE, "scene not committed");
    // preorder array + "left before right" == DFS leaf order; span-1 nodes list their primitive twice
    // in the reference (left == right), which this enumeration reproduces for comparison with the oracle
    if (!s->groups.empty()) {  // nested elements: the order the walk visits the primitives in (each once)
        const size_t n = std::min(cap, s->leaf_order_grouped.size());
        if (out && n) memcpy(out, s->leaf_order_grouped.data(), n * sizeof(int32_t));
        return (int64_t)s->leaf_order_grouped.size();
    }
    const int rc_nodes = host_nodes(s);
    if (rc_nodes != CR_OK) return rc_nodes;
    std::vector<int32_t> order;
    std::vector<uint32_t> stack;
    if (s->root != REF_MISS) stack.push_back(s->root);
    auto prim_index_of = [&](uint32_t ref) { return s->prim_of[ref_kind(ref)][ref_index(ref)]; };
    while (!stack.empty()) {
        const uint32_t ref = stack.back();
        stack.pop_back();
        if (ref_is_leaf(ref)) {
            order.push_back(prim_index_of(ref));
            continue;
        }
        const FlatNode& n = s->nodes[ref];
        if (n.right == REF_NONE) {
            order.push_back(prim_index_of(n.left));
            order.push_back(prim_index_of(n.left));
        } else {
            stack.push_back(n.right);
            stack.push_back(n.left);
        }
    }
    const size_t n = std::min(cap, order.size());
    if (out && n) memcpy(out, order.data(), n * sizeof(int32_t));
    return (int64_t)order.size();
}

int cr_trace_batch(CrScene* s, const double* rays, size_t n, double tmin, double tmax, int precision_flags, CrHit* out) {
    int rc = need_device(s);
    if (rc != CR_OK) return rc;
    if ((!rays || !out) && n) return fail(CR_ERR_INVALID, "null argument");
    const int precision = precision_flags & 0xff, reference_order = (precision_flags & CR_TRACE_REFERENCE_ORDER) ? 1 : 0;
    if ((precision != CR_PRECISION_F64 && precision != CR_PRECISION_F32) || (precision_flags & ~(0xff | CR_TRACE_REFERENCE_ORDER)))
        return fail(CR_ERR_INVALID, "bad precision");
    s->last_retried = 0;
    if (n == 0) return CR_OK;
    API_CUDA(cudaSetDevice(s->device));
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t bytes_in = n * 7 * sizeof(double), bytes_out = n * sizeof(CrHit), bytes_retry = n * sizeof(uint32_t);
    const size_t need = pad(bytes_in) + pad(bytes_out) + pad(bytes_retry) + 256;
    if (need > s->io_cap) {
        if (s->d_io) cudaFreeAsync(s->d_io, s->stream);
        s->d_io = nullptr;
        s->io_cap = 0;
        API_CUDA(cudaMallocAsync(&s->d_io, need, s->stream));
        s->io_cap = need;
    }
    char* base = static_cast<char*>(s->d_io);
    double* d_rays = reinterpret_cast<double*>(base);
    CrHit* d_out = reinterpret_cast<CrHit*>(base + pad(bytes_in));
    uint32_t* d_retry = reinterpret_cast<uint32_t*>(base + pad(bytes_in) + pad(bytes_out));
    uint32_t* d_cursor = reinterpret_cast<uint32_t*>(base + pad(bytes_in) + pad(bytes_out) + pad(bytes_retry));
    API_CUDA(cudaMemcpyAsync(d_rays, rays, bytes_in, cudaMemcpyHostToDevice, s->stream));
    std::string err;
    uint32_t retried = 0;
    rc = (precision == CR_PRECISION_F64)
             ? trace_batch_impl<double>(s->dev, d_rays, n, tmin, tmax, d_out, d_cursor, d_retry, reference_order, &retried, s->stream, err)
             : trace_batch_impl<float>(s->dev, d_rays, n, tmin, tmax, d_out, d_cursor, d_retry, reference_order, &retried, s->stream, err);
    if (rc != CR_OK) return fail(rc, err);
    API_CUDA(cudaMemcpyAsync(out, d_out, bytes_out, cudaMemcpyDeviceToHost, s->stream));
    API_CUDA(cudaStreamSynchronize(s->stream));
    s->last_retried = (int64_t)retried;
    return CR_OK;
}

int64_t cr_scene_last_retried(const CrScene* s) {
    if (!s) return fail(CR_ERR_INVALID, "null scene");
    return s->last_retried;
}

static int check_camera(const CrCamera* c) {
    if (!c) return fail(CR_ERR_INVALID, "null camera");
    if (c->image_width == 0 || c->image_height == 0) return fail(CR_ERR_INVALID, "empty image");
    // Camera::set_samples asserts s > 0 (camera/mod.rs:233-240)
    if (c->samples == 0) return fail(CR_ERR_INVALID, "The camera must have a positive number of samples. 0 is invalid.");
    if (c->n_from_keys > CR_MAX_CAM_KEYS || c->n_at_keys > CR_MAX_CAM_KEYS) return fail(CR_ERR_LIMIT, "too many camera keyframes");
    if (!(c->frame_rate > 0.0)) return fail(CR_ERR_INVALID, "frame_rate must be positive");
    if ((uint64_t)c->image_width * c->image_height > 0xFFFFFFFFull) return fail(CR_ERR_LIMIT, "image too large");
    return CR_OK;
}

int cr_render_device(CrScene* s, const CrCamera* cam, const CrRenderOpts* opts, void* d_out_rgb, void* d_out_rgb8,
                     void* cuda_stream, CrStats* stats) {
    int rc = need_device(s);
    if (rc != CR_OK) return rc;
    if ((rc = check_camera(cam)) != CR_OK) return rc;
    if (!opts) return fail(CR_ERR_INVALID, "null opts");
    if (opts->row_world > 1 && opts->row_rank >= opts->row_world) return fail(CR_ERR_INVALID, "row_rank >= row_world");
    API_CUDA(cudaSetDevice(s->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->stream;
    std::string err;
    std::lock_guard<std::mutex> device_lock(device_slot(s->device).mu);
    // device variant: rows of this rank are written PACKED ([rows_local][W][3]) so the result is the
    // NCCL gather send buffer as is
    const int packed = (opts->flags & CR_RENDER_GLOBAL_ROWS) ? 0 : 1;
    rc = (opts->precision == CR_PRECISION_F32)
             ? render_impl<float>(s->dev, device_workspace(s->device), *cam, *opts, d_out_rgb, d_out_rgb8, packed, st, stats, err)
             : render_impl<double>(s->dev, device_workspace(s->device), *cam, *opts, d_out_rgb, d_out_rgb8, packed, st, stats, err);
    if (rc != CR_OK) return fail(rc, err);
    return CR_OK;
}

int cr_render(CrScene* s, const CrCamera* cam, const CrRenderOpts* opts, double* out_rgb, uint8_t* out_rgb8, CrStats* stats) {
    int rc = need_device(s);
    if (rc != CR_OK) return rc;
    if ((rc = check_camera(cam)) != CR_OK) return rc;
    if (!opts) return fail(CR_ERR_INVALID, "null opts");
    if (opts->row_world > 1 && opts->row_rank >= opts->row_world) return fail(CR_ERR_INVALID, "row_rank >= row_world");
    API_CUDA(cudaSetDevice(s->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    const size_t npix = (size_t)cam->image_width * cam->image_height;
    if (npix > s->out_cap) {
        if (s->d_out_rgb) cudaFreeAsync(s->d_out_rgb, s->stream);
        if (s->d_out_rgb8) cudaFreeAsync(s->d_out_rgb8, s->stream);
        s->d_out_rgb = s->d_out_rgb8 = nullptr;
        s->out_cap = 0;
        API_CUDA(cudaMallocAsync(&s->d_out_rgb, npix * 3 * sizeof(double), s->stream));
        API_CUDA(cudaMallocAsync(&s->d_out_rgb8, npix * 3, s->stream));
        s->out_cap = npix;
    }
    std::lock_guard<std::mutex> device_lock(device_slot(s->device).mu);
    const bool sharded = opts->row_world > 1;
    if (sharded) {
        // untouched rows must stay untouched on the host: start from the caller's buffers
        if (out_rgb) API_CUDA(cudaMemcpyAsync(s->d_out_rgb, out_rgb, npix * 3 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (out_rgb8) API_CUDA(cudaMemcpyAsync(s->d_out_rgb8, out_rgb8, npix * 3, cudaMemcpyHostToDevice, s->stream));
    }
    std::string err;
    rc = (opts->precision == CR_PRECISION_F32)
             ? render_impl<float>(s->dev, device_workspace(s->device), *cam, *opts, out_rgb ? s->d_out_rgb : nullptr, out_rgb8 ? s->d_out_rgb8 : nullptr, 0,
                                  s->stream, stats, err)
             : render_impl<double>(s->dev, device_workspace(s->device), *cam, *opts, out_rgb ? s->d_out_rgb : nullptr, out_rgb8 ? s->d_out_rgb8 : nullptr,
                                   0, s->stream, stats, err);
    if (rc != CR_OK) return fail(rc, err);
    cudaEvent_t a, b;
    API_CUDA(cudaEventCreate(&a));
    API_CUDA(cudaEventCreate(&b));
    API_CUDA(cudaEventRecord(a, s->stream));
    if (out_rgb) API_CUDA(cudaMemcpyAsync(out_rgb, s->d_out_rgb, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out_rgb8) API_CUDA(cudaMemcpyAsync(out_rgb8, s->d_out_rgb8, npix * 3, cudaMemcpyDeviceToHost, s->stream));
    API_CUDA(cudaEventRecord(b, s->stream));
    API_CUDA(cudaEventSynchronize(b));
    if (stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        stats->ms_d2h = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return CR_OK;
}

// ---- scene export file: "CRSCENE1", then counted little-endian sections of the staging arrays ----------
extern "C++" {
namespace {
struct FileWriter {
    FILE* f;
    bool ok = true;
    void raw(const void* p, size_t n) { ok = ok && (n == 0 || fwrite(p, 1, n, f) == n); }
    void u64(uint64_t v) { raw(&v, 8); }
    template <typename T, typename A>
    void vec(const std::vector<T, A>& v) {
        u64(v.size());
        raw(v.data(), v.size() * sizeof(T));
    }
};
struct FileReader {
    FILE* f;
    bool ok = true;
    void raw(void* p, size_t n) { ok = ok && (n == 0 || fread(p, 1, n, f) == n); }
    uint64_t u64() {
        uint64_t v = 0;
        raw(&v, 8);
        return v;
    }
    template <typename T, typename A>
    void vec(std::vector<T, A>& v, uint64_t limit) {
        const uint64_t n = u64();
        if (!ok || n > limit) {
            ok = false;
            return;
        }
        v.resize((size_t)n);
        raw(v.data(), (size_t)n * sizeof(T));
    }
};
struct SavedElement {  // Element without the box (recomputed on load with the reference's arithmetic)
    uint32_t kind, idx, hide, pad;
};
}  // namespace
}  // extern "C++"

extern "C" int cr_scene_save(const CrScene* s, const char* path) {
    if (!s || !path) return fail(CR_ERR_INVALID, "null argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(CR_ERR_INVALID, std::string("cannot open ") + path);
    FileWriter w{f};
    w.raw("CRSCENE1", 8);
    std::vector<SavedElement> el(s->elements.size());
    for (size_t i = 0; i < el.size(); ++i) el[i] = {s->elements[i].kind, s->elements[i].idx, s->elements[i].hide ? 1u : 0u, 0u};
    w.vec(el);
    w.vec(s->spheres);
    w.vec(s->tris);
    w.vec(s->quads);
    for (int k = 0; k < 3; ++k) {
        w.vec(s->mat_of[k]);
        w.vec(s->obj_of[k]);
    }
    w.vec(s->mats);
    w.vec(s->texs);
    w.u64(s->images.size());
    for (const HostImage& im : s->images) {
        w.u64((uint64_t)im.w);
        w.u64((uint64_t)im.h);
        w.vec(im.rgb);
    }
    w.u64((uint64_t)(int64_t)s->sky_kind);
    w.u64((uint64_t)(int64_t)s->sky_image);
    w.u64((uint64_t)s->bvh_builder);
    w.u64(s->anim.size());
    for (const auto& kv : s->anim) {
        w.u64(kv.first);
        w.vec(kv.second);
    }
    // nested elements (optional trailing section: files without groups end here)
    if (!s->groups.empty()) {
        w.raw("GROUPS01", 8);
        w.u64(s->groups.size());
        for (const CrScene::Group& g : s->groups) {
            w.u64((uint64_t)(int64_t)g.kind);
            w.u64((uint64_t)(int64_t)g.parent);
            w.vec(g.members);
        }
        w.vec(s->top);
    }
    const bool ok = (fclose(f) == 0) && w.ok;
    return ok ? CR_OK : fail(CR_ERR_INVALID, std::string("write failed: ") + path);
}

extern "C" CrScene* cr_scene_load(const char* path, int device) {
    if (!path) {
        g_err = "cr_scene_load: null path";
        return nullptr;
    }
    FILE* f = fopen(path, "rb");
    if (!f) {
        g_err = std::string("cr_scene_load: cannot open ") + path;
        return nullptr;
    }
    CrScene* s = cr_scene_create(device);
    if (!s) {
        fclose(f);
        return nullptr;
    }
    FileReader r{f};
    char magic[8] = {0};
    r.raw(magic, 8);
    const uint64_t LIM = (uint64_t)REF_MAX_INDEX * 9;
    std::vector<SavedElement> el;
    if (r.ok && memcmp(magic, "CRSCENE1", 8) == 0) {
        r.vec(el, REF_MAX_INDEX);
        r.vec(s->spheres, LIM);
        r.vec(s->tris, LIM);
        r.vec(s->quads, LIM);
        for (int k = 0; k < 3; ++k) {
            r.vec(s->mat_of[k], REF_MAX_INDEX);
            r.vec(s->obj_of[k], REF_MAX_INDEX);
        }
        r.vec(s->mats, 1u << 24);
        r.vec(s->texs, 1u << 24);
        const uint64_t n_img = r.u64();
        for (uint64_t i = 0; r.ok && i < n_img && i < (1u << 20); ++i) {
            HostImage im;
            im.w = (int)r.u64();
            im.h = (int)r.u64();
            r.vec(im.rgb, 1ull << 34);
            if (r.ok && (im.w <= 0 || im.h <= 0 || im.rgb.size() != (size_t)im.w * im.h * 3)) r.ok = false;
            s->images.push_back(std::move(im));
        }
        s->sky_kind = (int)(int64_t)r.u64();
        s->sky_image = (int)(int64_t)r.u64();
        s->bvh_builder = (int)r.u64();
        const uint64_t n_anim = r.u64();
        for (uint64_t i = 0; r.ok && i < n_anim && i < LIM; ++i) {
            const uint64_t key = r.u64();
            r.vec(s->anim[key], 1u << 20);
        }
        char tag[8] = {0};
        if (r.ok && fread(tag, 1, 8, f) == 8) {  // optional: nested elements
            if (memcmp(tag, "GROUPS01", 8) != 0) r.ok = false;
            const uint64_t n_groups = r.ok ? r.u64() : 0;
            if (n_groups > 0x7FFFFFFEull) r.ok = false;
            for (uint64_t g = 0; r.ok && g < n_groups; ++g) {
                CrScene::Group grp;
                grp.kind = (int32_t)(int64_t)r.u64();
                grp.parent = (int32_t)(int64_t)r.u64();
                r.vec(grp.members, (uint64_t)REF_MAX_INDEX * 2);
                if (grp.kind != CR_GROUP_HITLIST && grp.kind != CR_GROUP_BVH) r.ok = false;
                if (grp.parent < -1 || grp.parent >= (int64_t)g) r.ok = false;
                s->groups.push_back(std::move(grp));
            }
            r.vec(s->top, (uint64_t)REF_MAX_INDEX * 2);
        }
    } else {
        r.ok = false;
    }
    fclose(f);
    // consistency of what was read, then the per-primitive tables the add calls would have built
    const size_t n_of[3] = {s->spheres.size() / 4, s->tris.size() / 9, s->quads.size() / 9};
    bool ok = r.ok && s->spheres.size() % 4 == 0 && s->tris.size() % 9 == 0 && s->quads.size() % 9 == 0 &&
              el.size() == n_of[0] + n_of[1] + n_of[2] && s->bvh_builder >= CR_BVH_AUTO && s->bvh_builder <= CR_BVH_DEVICE;
    for (int k = 0; ok && k < 3; ++k) ok = s->mat_of[k].size() == n_of[k] && s->obj_of[k].size() == n_of[k];
    if (ok) {
        for (int k = 0; k < 3; ++k) s->prim_of[k].assign(n_of[k], -1);
        s->elements.resize(el.size());
        for (size_t i = 0; ok && i < el.size(); ++i) {
            const SavedElement& e = el[i];
            if (e.kind > CR_PRIM_QUAD || e.idx >= n_of[e.kind] || s->prim_of[e.kind][e.idx] != -1) {
                ok = false;
                break;
            }
            const HostVec<double>& store = e.kind == CR_PRIM_SPHERE ? s->spheres : (e.kind == CR_PRIM_TRIANGLE ? s->tris : s->quads);
            const size_t stride = e.kind == CR_PRIM_SPHERE ? 4 : 9;
            for (size_t c = 0; c < stride; ++c)
                if (!std::isfinite(store[stride * e.idx + c])) ok = false;
            Element& dst = s->elements[i];
            dst.kind = e.kind;
            dst.idx = e.idx;
            dst.hide = e.hide != 0;
            dst.box = prim_box(e.kind, &store[stride * e.idx]);
            s->prim_of[e.kind][e.idx] = (int32_t)i;
        }
    }
    if (ok && !s->groups.empty()) {
        // every element and every group is a member of exactly one list, nested groups come after their parent
        std::vector<char> seen_e(s->elements.size(), 0), seen_g(s->groups.size(), 0);
        auto check = [&](const std::vector<uint32_t>& members, int64_t owner) {
            for (uint32_t m : members) {
                if (m & GROUP_MEMBER) {
                    const uint32_t g = m & ~GROUP_MEMBER;
                    if (g >= s->groups.size() || seen_g[g] || (int64_t)g <= owner || s->groups[g].parent != (int32_t)owner) return false;
                    seen_g[g] = 1;
                } else {
                    if (m >= s->elements.size() || seen_e[m]) return false;
                    seen_e[m] = 1;
                }
            }
            return true;
        };
        ok = check(s->top, -1);
        for (size_t g = 0; ok && g < s->groups.size(); ++g) ok = check(s->groups[g].members, (int64_t)g);
        for (size_t i = 0; ok && i < seen_e.size(); ++i) ok = seen_e[i] != 0;
        for (size_t g = 0; ok && g < seen_g.size(); ++g) ok = seen_g[g] != 0;
    }
    if (!ok) {
        cr_scene_destroy(s);
        g_err = std::string("cr_scene_load: ") + path + " is not a valid scene file";
        return nullptr;
    }
    if (cr_scene_commit(s) != CR_OK) {
        const std::string keep = g_err;
        cr_scene_destroy(s);
        g_err = keep;
        return nullptr;
    }
    return s;
}

// ---- multi-GPU behind the boundary (SURVEY 8b cr_init(devices, n), 8e) -------------------------------------------
extern "C" CrScene* cr_scene_replicate(const CrScene* src, int device) {
    if (!src) {
        g_err = "cr_scene_replicate: null scene";
        return nullptr;
    }
    CrScene* s = cr_scene_create(device);
    if (!s) return nullptr;
    s->elements = src->elements;
    s->spheres = src->spheres;
    s->tris = src->tris;
    s->quads = src->quads;
    for (int k = 0; k < 3; ++k) {
        s->mat_of[k] = src->mat_of[k];
        s->obj_of[k] = src->obj_of[k];
        s->prim_of[k] = src->prim_of[k];
    }
    s->mats = src->mats;
    s->texs = src->texs;
    for (const HostImage& im : src->images) {
        HostImage c;
        c.w = im.w;
        c.h = im.h;
        c.rgb = im.rgb;
        s->images.push_back(std::move(c));
    }
    s->sky_kind = src->sky_kind;
    s->sky_image = src->sky_image;
    s->anim = src->anim;
    s->bvh_builder = src->bvh_builder;
    s->groups = src->groups;
    s->top = src->top;
    if (cr_scene_commit(s) != CR_OK) {
        const std::string keep = g_err;
        cr_scene_destroy(s);
        g_err = keep;
        return nullptr;
    }
    return s;
}

extern "C" int cr_render_multi(CrScene* const* replicas, int n, const CrCamera* cam, const CrRenderOpts* opts, double* out_rgb,
                               uint8_t* out_rgb8, CrStats* stats) {
    if (!replicas || n < 1 || n > 64) return fail(CR_ERR_INVALID, "cr_render_multi: 1..64 replicas");
    int rc;
    for (int i = 0; i < n; ++i) {
        if ((rc = need_device(replicas[i])) != CR_OK) return rc;
        for (int j = 0; j < i; ++j)
            if (replicas[j]->device == replicas[i]->device) return fail(CR_ERR_INVALID, "cr_render_multi: two replicas on one device");
    }
    if ((rc = check_camera(cam)) != CR_OK) return rc;
    if (!opts) return fail(CR_ERR_INVALID, "null opts");
    CrScene* s0 = replicas[0];
    const int dev0 = s0->device;
    const size_t npix = (size_t)cam->image_width * cam->image_height;
    API_CUDA(cudaSetDevice(dev0));
    if (npix > s0->multi_cap) {
        if (s0->d_multi_rgb) cudaFree(s0->d_multi_rgb);
        if (s0->d_multi_rgb8) cudaFree(s0->d_multi_rgb8);
        s0->d_multi_rgb = s0->d_multi_rgb8 = nullptr;
        s0->multi_cap = 0;
        API_CUDA(cudaMalloc(&s0->d_multi_rgb, npix * 3 * sizeof(double)));
        API_CUDA(cudaMalloc(&s0->d_multi_rgb8, npix * 3));
        s0->multi_cap = npix;
    }
    // can every other device store into device 0's memory?  (NVLink / NVSwitch: yes)
    std::vector<char> direct((size_t)n, 1);
    for (int i = 1; i < n; ++i) {
        int can = 0;
        API_CUDA(cudaDeviceCanAccessPeer(&can, replicas[i]->device, dev0));
        if (can) {
            API_CUDA(cudaSetDevice(replicas[i]->device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(dev0, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
            cudaGetLastError();
        }
        direct[(size_t)i] = (char)can;
    }
    const uint32_t W = cam->image_width, H = cam->image_height;
    const uint32_t block = opts->row_block == 0 ? 8 : opts->row_block;
    std::vector<int> rcs((size_t)n, CR_OK);
    std::vector<std::string> errs((size_t)n);
    std::vector<CrStats> st((size_t)n);
    auto work = [&](int i) {
        CrScene* s = replicas[i];
        std::string& err = errs[(size_t)i];
        auto cuda_ok = [&](cudaError_t e, const char* what) {
            if (e == cudaSuccess) return true;
            rcs[(size_t)i] = CR_ERR_CUDA;
            err = std::string(what) + ": " + cudaGetErrorString(e);
            return false;
        };
        if (!cuda_ok(cudaSetDevice(s->device), "cudaSetDevice")) return;
        CrRenderOpts o = *opts;
        o.row_block = block;
        o.row_rank = (uint32_t)i;
        o.row_world = (uint32_t)n;
        memset(&st[(size_t)i], 0, sizeof(CrStats));
        void* d_rgb = out_rgb ? s0->d_multi_rgb : nullptr;
        void* d_rgb8 = out_rgb8 ? s0->d_multi_rgb8 : nullptr;
        int packed = 0;
        if (!direct[(size_t)i]) {
            // no peer mapping: render packed rows locally, then copy each row block to its place on device 0
            uint32_t rows_local = 0;
            for (uint32_t j = 0; j < H; ++j)
                if ((j / block) % (uint32_t)n == (uint32_t)i) ++rows_local;
            const size_t lp = (size_t)rows_local * W;
            if (lp > s->out_cap) {
                if (s->d_out_rgb) cudaFreeAsync(s->d_out_rgb, s->stream);
                if (s->d_out_rgb8) cudaFreeAsync(s->d_out_rgb8, s->stream);
                s->d_out_rgb = s->d_out_rgb8 = nullptr;
                s->out_cap = 0;
                if (!cuda_ok(cudaMallocAsync(&s->d_out_rgb, lp * 3 * sizeof(double), s->stream), "cudaMallocAsync")) return;
                if (!cuda_ok(cudaMallocAsync(&s->d_out_rgb8, lp * 3, s->stream), "cudaMallocAsync")) return;
                s->out_cap = lp;
            }
            d_rgb = out_rgb ? s->d_out_rgb : nullptr;
            d_rgb8 = out_rgb8 ? s->d_out_rgb8 : nullptr;
            packed = 1;
        }
        {
            std::lock_guard<std::mutex> device_lock(device_slot(s->device).mu);
            rcs[(size_t)i] = (o.precision == CR_PRECISION_F32)
                                 ? render_impl<float>(s->dev, device_workspace(s->device), *cam, o, d_rgb, d_rgb8, packed, s->stream, &st[(size_t)i], err)
                                 : render_impl<double>(s->dev, device_workspace(s->device), *cam, o, d_rgb, d_rgb8, packed, s->stream, &st[(size_t)i], err);
        }
        if (rcs[(size_t)i] != CR_OK || !packed) return;
        uint32_t lrow = 0;
        for (uint32_t j0 = 0; j0 < H; j0 += block) {
            if ((j0 / block) % (uint32_t)n != (uint32_t)i) continue;
            const uint32_t rows = std::min(block, H - j0);
            const size_t src = (size_t)lrow * W * 3, dst = (size_t)j0 * W * 3, cnt = (size_t)rows * W * 3;
            if (out_rgb && !cuda_ok(cudaMemcpyPeerAsync(static_cast<double*>(s0->d_multi_rgb) + dst, dev0, static_cast<double*>(d_rgb) + src,
                                                        s->device, cnt * sizeof(double), s->stream), "cudaMemcpyPeerAsync")) return;
            if (out_rgb8 && !cuda_ok(cudaMemcpyPeerAsync(static_cast<uint8_t*>(s0->d_multi_rgb8) + dst, dev0, static_cast<uint8_t*>(d_rgb8) + src,
                                                         s->device, cnt, s->stream), "cudaMemcpyPeerAsync")) return;
            lrow += rows;
        }
        cuda_ok(cudaStreamSynchronize(s->stream), "cudaStreamSynchronize");
    };
    {
        std::vector<std::thread> th;
        for (int i = 1; i < n; ++i) th.emplace_back(work, i);
        work(0);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < n; ++i)
        if (rcs[(size_t)i] != CR_OK) return fail(rcs[(size_t)i], "replica " + std::to_string(i) + ": " + errs[(size_t)i]);
    // every replica's stream has drained, so every peer store has landed: one device-to-host copy of the assembled image
    API_CUDA(cudaSetDevice(dev0));
    cudaEvent_t a, b;
    API_CUDA(cudaEventCreate(&a));
    API_CUDA(cudaEventCreate(&b));
    API_CUDA(cudaEventRecord(a, s0->stream));
    if (out_rgb) API_CUDA(cudaMemcpyAsync(out_rgb, s0->d_multi_rgb, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost, s0->stream));
    if (out_rgb8) API_CUDA(cudaMemcpyAsync(out_rgb8, s0->d_multi_rgb8, npix * 3, cudaMemcpyDeviceToHost, s0->stream));
    API_CUDA(cudaEventRecord(b, s0->stream));
    API_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    st[0].ms_d2h = ms;
    if (stats)
        for (int i = 0; i < n; ++i) stats[i] = st[(size_t)i];
    return CR_OK;
}

extern "C" int cr_shared_buffer_create(int device, size_t bytes, void** dptr, unsigned char handle[64]) {
    if (!dptr || !handle || bytes == 0) return fail(CR_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI ships the IPC handle as 64 bytes");
    const DeviceInfo& di = device_info(device);
    if (di.state != 1) return fail(CR_ERR_NO_DEVICE, di.why);
    API_CUDA(cudaSetDevice(device));
    void* p = nullptr;
    API_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(CR_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle, &h, 64);
    *dptr = p;
    return CR_OK;
}
extern "C" int cr_shared_buffer_open(int device, const unsigned char handle[64], void** dptr) {
    if (!dptr || !handle) return fail(CR_ERR_INVALID, "null argument");
    const DeviceInfo& di = device_info(device);
    if (di.state != 1) return fail(CR_ERR_NO_DEVICE, di.why);
    API_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    API_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr = p;
    return CR_OK;
}
extern "C" int cr_shared_buffer_close(int device, void* dptr, int owner) {
    if (!dptr) return CR_OK;
    const DeviceInfo& di = device_info(device);
    if (di.state != 1) return fail(CR_ERR_NO_DEVICE, di.why);
    API_CUDA(cudaSetDevice(device));
    if (owner) API_CUDA(cudaFree(dptr));
    else API_CUDA(cudaIpcCloseMemHandle(dptr));
    return CR_OK;
}

// ---- Camera::render's file tail (camera/mod.rs:275-311) and Scene::render_movie's frame loop ----------
namespace {

struct DecimalLut {  // "0".."255" followed by the separator the caller appends
    char txt[256][4];
    uint8_t len[256];
    DecimalLut() {
        for (int v = 0; v < 256; ++v) len[v] = (uint8_t)snprintf(txt[v], 4, "%d", v);
    }
};

// "{r} {g} {b}\n" for pixels [p0, p1) appended to out (Display for Color, utils.rs:436)
void format_p3_rows(const uint8_t* rgb8, size_t p0, size_t p1, std::string& out) {
    static const DecimalLut lut;
    out.resize((p1 - p0) * 12);
    char* w = &out[0];
    for (size_t p = p0; p < p1; ++p) {
        const uint8_t* c = rgb8 + 3 * p;
        for (int k = 0; k < 3; ++k) {
            const uint8_t v = c[k];
            memcpy(w, lut.txt[v], 3);
            w += lut.len[v];
            *w++ = (k == 2) ? '\n' : ' ';
        }
    }
    out.resize((size_t)(w - &out[0]));
}

// EXTENSION: 8-bit RGB PNG.  Rows with filter type 0, split into bands that host threads deflate
// independently (zlib level 1; every band but the last ends with a sync flush, so the concatenation is one valid zlib
// stream: header from the first band, Adler-32 of the whole image appended), chunks IHDR / IDAT / IEND with CRC-32.
void png_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
    auto be32 = [&](uint32_t v) {
        out.push_back((uint8_t)(v >> 24)); out.push_back((uint8_t)(v >> 16)); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v);
    };
    be32((uint32_t)n);
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    be32((uint32_t)crc32(0L, out.data() + at, (uInt)(n + 4)));
}
int write_png_file(const char* path, const uint8_t* rgb8, uint32_t w, uint32_t h) {
    if (w == 0 || h == 0) return fail(CR_ERR_INVALID, "empty image");
    const size_t stride = (size_t)w * 3;
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1 : (nt > 8 ? 8 : nt);
    if ((size_t)w * h < 65536 || h < nt) nt = 1;
    std::vector<std::vector<uint8_t>> band(nt);
    std::vector<int> rc(nt, Z_OK);
    std::vector<uLong> adler(nt, 1), band_len(nt, 0);
    auto work = [&](unsigned t) {
        const uint32_t r0 = (uint32_t)((uint64_t)h * t / nt), r1 = (uint32_t)((uint64_t)h * (t + 1) / nt);
        std::vector<uint8_t> raw((size_t)(r1 - r0) * (stride + 1));
        for (uint32_t r = r0; r < r1; ++r) {
            uint8_t* dst = raw.data() + (size_t)(r - r0) * (stride + 1);
            dst[0] = 0;  // filter type 0 (None)
            memcpy(dst + 1, rgb8 + (size_t)r * stride, stride);
        }
        adler[t] = adler32(1L, raw.data(), (uInt)raw.size());
        band_len[t] = (uLong)raw.size();
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, 1, Z_DEFLATED, t == 0 ? 15 : -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {  // band 0 carries the zlib header
            rc[t] = Z_MEM_ERROR;
            return;
        }
        band[t].resize(deflateBound(&zs, (uLong)raw.size()) + 64);
        zs.next_in = raw.data();
        zs.avail_in = (uInt)raw.size();
        zs.next_out = band[t].data();
        zs.avail_out = (uInt)band[t].size();
        const int r = deflate(&zs, t + 1 == nt ? Z_FINISH : Z_FULL_FLUSH);
        if (r != (t + 1 == nt ? Z_STREAM_END : Z_OK)) rc[t] = Z_BUF_ERROR;
        size_t produced = band[t].size() - zs.avail_out;
        if (t == 0 && nt == 1 && produced >= 4) produced -= 4;  // single band: zlib appended its own Adler-32; ours follows below
        band[t].resize(produced);
        deflateEnd(&zs);
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    for (unsigned t = 0; t < nt; ++t)
        if (rc[t] != Z_OK) return fail(CR_ERR_INVALID, "PNG: deflate failed");
    uLong ad = adler[0];
    for (unsigned t = 1; t < nt; ++t) ad = adler32_combine(ad, adler[t], (z_off_t)band_len[t]);
    std::vector<uint8_t> idat;
    for (unsigned t = 0; t < nt; ++t) idat.insert(idat.end(), band[t].begin(), band[t].end());
    idat.push_back((uint8_t)(ad >> 24)); idat.push_back((uint8_t)(ad >> 16)); idat.push_back((uint8_t)(ad >> 8)); idat.push_back((uint8_t)ad);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h,
                        8, 2, 0, 0, 0};  // 8 bits per sample, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    png_chunk(out, "IHDR", ihdr, 13);
    png_chunk(out, "IDAT", idat.data(), idat.size());
    png_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(CR_ERR_INVALID, std::string("cannot open ") + path);
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = (fclose(f) == 0) && ok;
    return ok ? CR_OK : fail(CR_ERR_INVALID, std::string("write failed: ") + path);
}

int write_ppm_file(const char* path, const uint8_t* rgb8, uint32_t w, uint32_t h, int format) {
    if (!path || (!rgb8 && w && h)) return fail(CR_ERR_INVALID, "null argument");
    if (format != CR_PPM_P3 && format != CR_PPM_P6 && format != CR_PNG) return fail(CR_ERR_INVALID, "bad image format");
    if (format == CR_PNG) return write_png_file(path, rgb8, w, h);
    FILE* f = fopen(path, "wb");  // create or truncate, camera/mod.rs:275-280
    if (!f) return fail(CR_ERR_INVALID, std::string("cannot open ") + path);
    const size_t npix = (size_t)w * h;
    bool ok = fprintf(f, "%s\n%u %u\n255\n", format == CR_PPM_P3 ? "P3" : "P6", w, h) > 0;
    if (format == CR_PPM_P6) {
        ok = ok && fwrite(rgb8, 1, npix * 3, f) == npix * 3;
    } else {
        unsigned nt = std::thread::hardware_concurrency();
        nt = nt == 0 ? 1 : (nt > 8 ? 8 : nt);
        if (npix < 65536) nt = 1;
        std::vector<std::string> chunk(nt);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) {
            const size_t p0 = npix * t / nt, p1 = npix * (t + 1) / nt;
            if (t + 1 == nt) format_p3_rows(rgb8, p0, p1, chunk[t]);
            else th.emplace_back(format_p3_rows, rgb8, p0, p1, std::ref(chunk[t]));
        }
        for (auto& x : th) x.join();
        for (unsigned t = 0; t < nt && ok; ++t) ok = fwrite(chunk[t].data(), 1, chunk[t].size(), f) == chunk[t].size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? CR_OK : fail(CR_ERR_INVALID, std::string("write failed: ") + path);
}

}  // namespace

extern "C" int cr_write_ppm(const char* path, const uint8_t* rgb8, uint32_t w, uint32_t h, int format) {
    return write_ppm_file(path, rgb8, w, h, format);
}

extern "C" int cr_render_to_file(CrScene* s, const CrCamera* cam, const CrRenderOpts* opts, const char* path, int format, CrStats* stats) {
    if (!path) return fail(CR_ERR_INVALID, "null path");
    if (format != CR_PPM_P3 && format != CR_PPM_P6 && format != CR_PNG) return fail(CR_ERR_INVALID, "bad image format");
    if (opts && opts->row_world > 1) return fail(CR_ERR_INVALID, "cr_render_to_file renders whole images (row_world <= 1)");
    int rc = check_camera(cam);
    if (rc != CR_OK) return rc;
    std::vector<uint8_t> img((size_t)cam->image_width * cam->image_height * 3);
    rc = cr_render(s, cam, opts, nullptr, img.data(), stats);
    if (rc != CR_OK) return rc;
    return write_ppm_file(path, img.data(), cam->image_width, cam->image_height, format);
}

extern "C" int cr_render_frames(CrScene* s, const CrCamera* cam, const CrRenderOpts* opts, uint32_t first, uint32_t stride,
                                uint32_t n_frames, const char* dir, uint32_t digits, int format, CrStats* stats) {
    int rc = need_device(s);
    if (rc != CR_OK) return rc;
    if ((rc = check_camera(cam)) != CR_OK) return rc;
    if (!opts || !dir) return fail(CR_ERR_INVALID, "null argument");
    if (stride == 0) return fail(CR_ERR_INVALID, "stride must be positive");
    if (opts->row_world > 1) return fail(CR_ERR_INVALID, "cr_render_frames shards whole frames (row_world <= 1)");
    if (format != CR_PPM_P3 && format != CR_PPM_P6 && format != CR_PNG) return fail(CR_ERR_INVALID, "bad image format");
    API_CUDA(cudaSetDevice(s->device));
    const uint32_t W = cam->image_width, H = cam->image_height;
    const size_t bytes = (size_t)W * H * 3;
    constexpr int NBUF = 3;
    uint8_t* dbuf[NBUF] = {nullptr, nullptr, nullptr};
    uint8_t* hbuf[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t copied[NBUF];
    cudaStream_t copy_stream = nullptr;
    bool busy[NBUF] = {false, false, false};
    struct Job {
        int buf;
        uint32_t frame;
    };
    std::deque<Job> jobs;
    std::mutex mu;
    std::condition_variable cv;
    bool closing = false;
    int write_rc = CR_OK;
    std::string write_err;

    auto cleanup = [&]() {
        for (int b = 0; b < NBUF; ++b) {
            if (dbuf[b]) cudaFreeAsync(dbuf[b], s->stream);
            if (hbuf[b]) cudaFreeHost(hbuf[b]);
            if (copied[b]) cudaEventDestroy(copied[b]);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        cudaGetLastError();  // a failed allocation above must not surface in the next call's error check
    };
    for (int b = 0; b < NBUF; ++b) copied[b] = nullptr;
    for (int b = 0; b < NBUF; ++b) {
        if (cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming) != cudaSuccess ||
            cudaMallocAsync((void**)&dbuf[b], bytes, s->stream) != cudaSuccess ||
            cudaHostAlloc((void**)&hbuf[b], bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            cleanup();
            return fail(CR_ERR_CUDA, "cr_render_frames: buffer allocation failed");
        }
    }
    if (cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cleanup();
        return fail(CR_ERR_CUDA, "cr_render_frames: cudaStreamCreate failed");
    }

    // writer: waits for a frame's bytes to land in pinned memory, formats and writes the file while the
    // GPU traces the next frame
    const int device = s->device;
    std::thread writer([&]() {
        cudaSetDevice(device);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return closing || !jobs.empty(); });
                if (jobs.empty()) return;
                j = jobs.front();
                jobs.pop_front();
            }
            int r = CR_OK;
            std::string e;
            if (cudaEventSynchronize(copied[j.buf]) != cudaSuccess) {
                r = CR_ERR_CUDA;
                e = "cr_render_frames: device to host copy failed";
            } else {
                char name[64];
                snprintf(name, sizeof(name), "/image%0*u.%s", (int)digits, j.frame, format == CR_PNG ? "png" : "ppm");  // scene/mod.rs:307-308
                r = write_ppm_file((std::string(dir) + name).c_str(), hbuf[j.buf], W, H, format);
                if (r != CR_OK) e = g_err;  // thread-local of the writer thread
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                busy[j.buf] = false;
                if (r != CR_OK && write_rc == CR_OK) {
                    write_rc = r;
                    write_err = e;
                }
            }
            cv.notify_all();
        }
    });

    std::string err;
    uint32_t k = 0;
    for (uint32_t frame = first; frame < n_frames && rc == CR_OK; frame += stride, ++k) {
        const int b = (int)(k % NBUF);
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !busy[b]; });
            if (write_rc != CR_OK) break;
            busy[b] = true;
        }
        CrCamera c = *cam;
        c.frame = cam->frame + frame;  // the loop counter names the file, Camera.frame advances from where it stood
        CrStats st;
        memset(&st, 0, sizeof(st));
        {
            std::lock_guard<std::mutex> device_lock(device_slot(s->device).mu);
            rc = (opts->precision == CR_PRECISION_F32)
                     ? render_impl<float>(s->dev, device_workspace(s->device), c, *opts, nullptr, dbuf[b], 0, s->stream, &st, err)
                     : render_impl<double>(s->dev, device_workspace(s->device), c, *opts, nullptr, dbuf[b], 0, s->stream, &st, err);
        }
        if (rc != CR_OK) {
            std::lock_guard<std::mutex> lk(mu);
            busy[b] = false;
            break;
        }
        if (stats) stats[k] = st;
        // render_impl returns after its stream has drained: the copy can start at once on the second stream
        if (cudaMemcpyAsync(hbuf[b], dbuf[b], bytes, cudaMemcpyDeviceToHost, copy_stream) != cudaSuccess ||
            cudaEventRecord(copied[b], copy_stream) != cudaSuccess) {
            rc = CR_ERR_CUDA;
            err = "cr_render_frames: cudaMemcpyAsync failed";
            std::lock_guard<std::mutex> lk(mu);
            busy[b] = false;
            break;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            jobs.push_back({b, frame});
        }
        cv.notify_all();
    }
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return jobs.empty() && !busy[0] && !busy[1] && !busy[2]; });
        closing = true;
    }
    cv.notify_all();
    writer.join();
    cudaStreamSynchronize(copy_stream);
    cleanup();
    if (rc != CR_OK) return fail(rc, err);
    if (write_rc != CR_OK) return fail(write_rc, write_err);
    return CR_OK;
}

// TransformTimeline::combine_and_compute for a camera point, timeline/mod.rs:233-263 (host copy of
// the device routine; used by the frame-sharding driver and by tests)
int cr_camera_point_at(const double init[3], const CrKeyframe* keys, size_t n, double t, double out[3]) {
    if (!init || !out || (!keys && n)) return fail(CR_ERR_INVALID, "null argument");
    double p[3] = {init[0], init[1], init[2]};
    for (size_t k = 0; k < n; ++k) {
        const double t0 = keys[k].t0, t1 = keys[k].t1;
        if (keys[k].axis < 0 || keys[k].axis > 2) return fail(CR_ERR_INVALID, "keyframe axis out of range");
        if (!((t > t1) || (t0 <= t && t <= t1))) continue;
        double s = (t - t0) / (t1 - t0);
        s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
        const double off = (keys[k].interp == CR_LERP) ? keys[k].delta * s : keys[k].delta;
        p[keys[k].axis] = off + p[keys[k].axis];
    }
    out[0] = p[0];
    out[1] = p[1];
    out[2] = p[2];
    return CR_OK;
}

// TransformTimeline::combine_and_compute for an object point, timeline/mod.rs:233-263: the routine the trace / shade
// kernels run at the ray's time (anim_eval, common.cuh), compiled here for the host (no FMA contraction)
int cr_anim_point_at(const double init[4], const CrAnimKey* keys, size_t n, double t, double out[4]) {
    if (!init || !out || (!keys && n)) return fail(CR_ERR_INVALID, "null argument");
    if (n > 0xFFFFFFFFull) return fail(CR_ERR_LIMIT, "too many keyframes");
    for (size_t i = 0; i < n; ++i) {
        if (keys[i].kind < 0 || keys[i].kind > 6) return fail(CR_ERR_INVALID, "keyframe kind out of range");
        if (keys[i].interp != CR_NERP && keys[i].interp != CR_LERP) return fail(CR_ERR_INVALID, "bad interpolation type");
    }
    double p[3] = {init[0], init[1], init[2]}, w = init[3];
    anim_eval<double>(keys, 0u, (uint32_t)n, t, p, w);
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = w;
    return CR_OK;
}

int cr_measure_fma_peak(int device, double* fp64_tflops, double* fp32_tflops) {
    if (!fp64_tflops || !fp32_tflops) return fail(CR_ERR_INVALID, "null argument");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        return fail(CR_ERR_NO_DEVICE, "no CUDA device");
    }
    API_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    API_CUDA(cudaGetDeviceProperties(&p, device));
    std::string err;
    int rc = measure_fma_peak(p.multiProcessorCount, fp64_tflops, fp32_tflops, err);
    if (rc != CR_OK) return fail(rc, err);
    return CR_OK;
}

void cr_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

}  // extern "C"
