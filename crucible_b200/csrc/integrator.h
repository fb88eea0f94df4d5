// integrator.h — interface between the C-ABI layer (api.cu) and the two precision-specific
// translation units (integrator_f64.cu compiled with -fmad=false, integrator_f32.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "common.cuh"

namespace crb {

// Type-erased device pointers of a committed scene; index 0 = f64 records, 1 = f32 records.
struct SceneDeviceData {
    void* nodes[2] = {nullptr, nullptr};
    void* spheres[2] = {nullptr, nullptr};
    void* tris[2] = {nullptr, nullptr};
    void* quads[2] = {nullptr, nullptr};
    PrimMeta* meta[3] = {nullptr, nullptr, nullptr};
    DevMaterial* mats = nullptr;
    DevTexture* texs = nullptr;
    DevImage* images = nullptr;
    CrAnimKey* anim_keys = nullptr;
    AnimTrack* sphere_track = nullptr;
    AnimTrack* tri_track = nullptr;
    uint32_t* tri_anim_slot = nullptr;
    double* tri_anim_verts = nullptr;
    void* fast_nodes = nullptr;  // search tree of the order-free engine (FastNodeRec[]), nullptr = none
    void* fast_prims = nullptr;  // its leaf-primitive table (uint2[])
    uint32_t n_fast_nodes = 0, n_fast_prims = 0;  // host-built search trees only (0: the tree was built on the device)
    uint32_t n_nodes = 0;
    uint32_t n_prims[3] = {0, 0, 0};  // spheres, triangles, quads
    int32_t sky_kind = CR_SKY_DEFAULT, sky_image = -1;
    int32_t clamp_colors = 1;
    float bmax = 0.f;            // largest |coordinate| of the root box (f32, rounded up)
    float bsmall = 0.f;          // bound of the nodes without BIGBOX_BIT
    double max_radiance = 1.0;  // largest per-sample colour component (1 unless the scene holds an Emissive)
    int num_sms = 148;
    int node_slice = 32;
    int strict_boxes = 0;  // nested elements (cr_scene_begin_group): see DevScene::strict_boxes
};

// Grow-only device scratch owned by the CrScene (path pool, queues, fixed-point framebuffer).
struct Workspace {
    void* ptr = nullptr;
    size_t bytes = 0;
    void* pinned = nullptr;  // 4 KiB of pinned host memory for control-block staging
    // trace-kernel variant chosen by timing, per scene signature (render_impl)
    uint64_t tuned_sig[16] = {0};
    int tuned_variant[16] = {0};
    int tuned_n = 0;
    int lookup_variant(uint64_t sig) const {
        for (int i = 0; i < tuned_n && i < 16; ++i)
            if (tuned_sig[i] == sig) return tuned_variant[i];
        return -1;
    }
    void store_variant(uint64_t sig, int v) {
        const int i = tuned_n % 16;  // small ring: a process renders a handful of distinct scenes
        tuned_sig[i] = sig;
        tuned_variant[i] = v;
        ++tuned_n;
        if (tuned_n >= 32) tuned_n = 16;
    }
    int ensure(size_t need, std::string& err) {
        if (need <= bytes) return CR_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&ptr, need);
        if (e != cudaSuccess) {
            err = std::string("cudaMalloc(workspace): ") + cudaGetErrorString(e);
            return CR_ERR_CUDA;
        }
        bytes = need;
        return CR_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        if (pinned) cudaFreeHost(pinned);
        ptr = nullptr;
        pinned = nullptr;
        bytes = 0;
    }
};

template <typename R>
int render_impl(const SceneDeviceData&, Workspace&, const CrCamera&, const CrRenderOpts&, void* d_out_rgb, void* d_out_rgb8,
                int packed, cudaStream_t, CrStats*, std::string& err);
template <typename R>
int trace_batch_impl(const SceneDeviceData&, const double* d_rays, size_t n, double tmin, double tmax, CrHit* d_out,
                     uint32_t* d_cursor, uint32_t* d_retry, int reference_order, uint32_t* h_retried, cudaStream_t, std::string& err);

// Device build of the order-free engine's search tree for large scenes (search_tree.cu)
int gpu_build_search_tree(const SceneDeviceData& d, cudaStream_t stream, uint32_t n_visible, void** d_fast_nodes, void** d_fast_prims,
                          uint32_t* depth, std::string& err);

// FMA micro-benchmarks (roofline denominators); defined in integrator_f32.cu
int measure_fma_peak(int num_sms, double* fp64_tflops, double* fp32_tflops, std::string& err);

}  // namespace crb
