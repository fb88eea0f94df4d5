// integrator.cuh — the wavefront path tracer: raygen -> persistent-thread trace -> material-sorted
// shade kernels with stream compaction -> resolve.  Everything is templated on R (double = faithful,
// float = fast) and instantiated by integrator_f64.cu (-fmad=false) and integrator_f32.cu.
//
// Replaces Camera::cast_ray / ray_color / average_samples (src/camera/ray_casting.rs:64-173) and the
// worker pool of src/camera/cpu_threading.rs:25-115.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>

#include "device_math.cuh"
#include "fast_trace.cuh"
#include "integrator.h"

namespace crb {

static constexpr int TRACE_BLOCK = 128;
static constexpr int SHADE_BLOCK = 128;

// ---- warp-aggregated append (one atomic per warp per destination) --------------------------------
static __device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool pred) {
    const uint32_t mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

// Enqueue every lane with q >= 0 into queue q: Q_COUNT ballots, then lanes 0..Q_COUNT-1 each issue the ONE
// atomic of "their" queue (distinct addresses, one instruction), bases come back by shuffle.
static __device__ __forceinline__ uint32_t warp_enqueue(uint32_t* counts, int q) {
    const int lane = threadIdx.x & 31;
    uint32_t my_mask = 0, lane_cnt = 0;
#pragma unroll
    for (int k = 0; k < Q_COUNT; ++k) {
        const uint32_t m = __ballot_sync(0xffffffffu, q == k);
        if (q == k) my_mask = m;
        if (lane == k) lane_cnt = (uint32_t)__popc(m);
    }
    uint32_t base = 0;
    if (lane_cnt) base = atomicAdd(counts + lane, lane_cnt);
    base = __shfl_sync(0xffffffffu, base, q < 0 ? 0 : q);
    return base + (uint32_t)__popc(my_mask & ((1u << lane) - 1u));
}

// ---- textures (src/textures/*.rs) -------------------------------------------------------------------
// image_texture.rs:23-32 and SkyboxImage::get_color (scene/mod.rs:37-45): clamp, flip v, truncate,
// clamp to the last texel (img_loader.rs:69-77), byte / 255.0 (img_loader.rs:40-42).  The fetch goes
// through a point-sampled CUDA texture object (uchar4), so the texel is exact and cached by the TEX unit.
template <typename R>
__device__ __forceinline__ V3<R> image_lookup(const DevImage& im, R u, R v) {
    u = rclamp(u, R(0), R(1));
    v = R(1) - rclamp(v, R(0), R(1));
    R fi = u * (R)im.w, fj = v * (R)im.h;
    // `as usize`: truncation toward zero, saturating, NaN -> 0
    long long i = (fi == fi) ? (long long)fi : 0;
    long long j = (fj == fj) ? (long long)fj : 0;
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (i > im.w - 1) i = im.w - 1;
    if (j > im.h - 1) j = im.h - 1;
    const uchar4 t = tex2D<uchar4>(im.tex, (float)i + 0.5f, (float)j + 0.5f);
    return {(R)t.x / R(255), (R)t.y / R(255), (R)t.z / R(255)};
}
template <typename R>
__device__ __forceinline__ int32_t floor_i32(R x) {  // `.floor() as i32` (saturating, NaN -> 0)
    R f = Num<R>::floor_(x);
    if (!(f == f)) return 0;
    if (f <= R(-2147483648.0)) return INT32_MIN;
    if (f >= R(2147483647.0)) return INT32_MAX;
    return (int32_t)f;
}
template <typename R>
__device__ __forceinline__ V3<R> tex_value(const DevScene<R>& sc, int tex, R u, R v, V3<R> p) {
    for (int nest = 0; nest < MAX_TEX_NEST; ++nest) {
        const DevTexture& t = sc.texs[tex];
        if (t.kind == CR_TEX_SOLID) return {(R)t.color[0], (R)t.color[1], (R)t.color[2]};  // solid_color.rs:25-27
        if (t.kind == CR_TEX_IMAGE) return image_lookup<R>(sc.images[t.image], u, v);
        // checker_texture.rs:39-51
        const R s = (R)t.inv_scale;
        const int32_t xi = floor_i32(s * p.x), yi = floor_i32(s * p.y), zi = floor_i32(s * p.z);
        const int32_t sum = (int32_t)((uint32_t)xi + (uint32_t)yi + (uint32_t)zi);
        tex = (sum % 2 == 0) ? t.even : t.odd;
    }
    return {R(0), R(0), R(0)};
}

// ---- sky, src/camera/ray_casting.rs:133-151 ---------------------------------------------------------
template <typename R>
__device__ __forceinline__ V3<R> sky_color(const DevScene<R>& sc, V3<R> d) {
    const bool cl = sc.clamp_colors != 0;
    if (sc.sky_kind == CR_SKY_SPHERICAL) {
        V3<R> ud = vunit(d);
        R theta = Num<R>::atan2_(ud.x, ud.z);
        R phi = Num<R>::asin_(ud.y);
        R u = (theta / (R(2) * Num<R>::pi())) + R(0.5);
        R v = (phi / Num<R>::pi()) + R(0.5);
        return image_lookup<R>(sc.images[sc.sky_image], u, v);
    }
    if (sc.sky_kind == CR_SKY_BLACK) return {R(0), R(0), R(0)};
    V3<R> ud = vunit(d);
    R a = R(0.5) * (ud.y + R(1));
    return col_add(col_scale(R(1) - a, V3<R>{R(1), R(1), R(1)}, cl), col_scale(a, V3<R>{R(0.5), R(0.7), R(1)}, cl), cl);
}

// ---- camera, src/camera/rendering_compute.rs + ray_casting.rs:77-104 -----------------------------
// TransformTimeline::combine_and_compute for a camera point (timeline/mod.rs:233-263)
template <typename R>
__device__ __forceinline__ V3<R> point_at(const double init[3], const CrKeyframe* keys, uint32_t n, R t) {
    R p[3] = {(R)init[0], (R)init[1], (R)init[2]};
    for (uint32_t k = 0; k < n; ++k) {
        const R t0 = (R)keys[k].t0, t1 = (R)keys[k].t1;
        if (!((t > t1) || (t0 <= t && t <= t1))) continue;
        R s = rclamp((t - t0) / (t1 - t0), R(0), R(1));
        R off = (keys[k].interp == CR_LERP) ? (R)keys[k].delta * s : (R)keys[k].delta;
        const int ax = keys[k].axis;
        p[ax] = off + p[ax];
    }
    return {p[0], p[1], p[2]};
}
// Camera basis at time t.  Every basis function of the reference (rendering_compute.rs:5-111) is a pure
// function of t, so evaluating each once gives the same bits as the reference's ~20 re-evaluations.
template <typename R>
struct CamBasis {
    V3<R> from, psl, pdu, pdv, du, dv;
};
template <typename R>
__device__ __forceinline__ CamBasis<R> camera_basis(const CrCamera& c, R t) {
    CamBasis<R> b;
    const V3<R> from = point_at<R>(c.look_from, c.from_keys, c.n_from_keys, t);
    const V3<R> at = point_at<R>(c.look_at, c.at_keys, c.n_at_keys, t);
    const V3<R> vup = {(R)c.vup[0], (R)c.vup[1], (R)c.vup[2]};
    const V3<R> w = vunit(vsub(from, at));                          // w_basis :88-93
    const V3<R> u = vunit(vcross(vup, w));                          // u_basis :77-80
    const V3<R> v = vcross(w, u);                                   // v_basis :82-85
    const V3<R> vu = vmul((R)c.viewport_width, u);                  // viewport_u :16-19
    const V3<R> vv = vmul((R)c.viewport_height, vneg(v));           // viewport_v :24-27
    b.pdu = vdiv(vu, (R)c.image_width);                             // pixel_delta_u :32-35
    b.pdv = vdiv(vv, (R)c.image_height);                            // pixel_delta_v :40-43
    const V3<R> ul = vsub(vsub(vsub(from, vmul((R)c.focus_dist, w)), vdiv(vu, R(2))), vdiv(vv, R(2)));  // :49-55
    b.psl = vadd(ul, vmul(R(0.5), vadd(b.pdu, b.pdv)));             // pixel_start_location :57-60
    b.du = vmul((R)c.defocus_radius, u);                            // defocus_disk_u/v :95-103
    b.dv = vmul((R)c.defocus_radius, v);
    b.from = from;
    return b;
}
template <typename R> __device__ __forceinline__ V3<R> ld3(const double* p) { return {(R)p[0], (R)p[1], (R)p[2]}; }
// One iteration of the sample loop, ray_casting.rs:82-104
template <typename R>
__device__ __forceinline__ void camera_sample(const DevCamera& cam, uint32_t i, uint32_t j, Rng<R>& g, V3<R>& ro, V3<R>& rd,
                                              R& tm) {
    const CrCamera& c = cam.c;
    const R current_time = (R)c.frame * (R(1) / (R)c.frame_rate);
    const R shutter_length = ((R)c.shutter_angle / R(360)) * (R(1) / (R)c.frame_rate);
    const R t = current_time + g.range(R(0), shutter_length);
    const CamBasis<R> b = camera_basis<R>(c, t);
    const R ox = g.next() - R(0.5);  // sample_square, camera/mod.rs:369-376
    const R oy = g.next() - R(0.5);
    const V3<R> ps = vadd(vadd(b.psl, vmul((R)i + ox, b.pdu)), vmul((R)j + oy, b.pdv));  // get_pixel_pos :64-68
    V3<R> orig = b.from;
    if (!(c.defocus_angle <= 0.0)) {                                // ray_casting.rs:96-100
        const V3<R> p = random_in_unit_disk(g);                     // defocus_disk_sample :105-110
        orig = vadd(vadd(b.from, vmul(p.x, b.du)), vmul(p.y, b.dv));
    }
    ro = orig;
    rd = vsub(ps, orig);
    tm = t;
}
// local (per-rank) row -> global row: rows j with (j / row_block) % row_world == row_rank, in order
__host__ __device__ inline uint32_t local_to_global_row(uint32_t lr, uint32_t block, uint32_t rank, uint32_t world) {
    if (world <= 1) return lr;
    const uint32_t tile = lr / block, within = lr % block;
    return (tile * world + rank) * block + within;
}

// ---- kernels --------------------------------------------------------------------------------------
// plan: single thread.  Closes iteration `side`->`nxt`: survivors are already in side nxt; decide how
// many camera samples top the pool up, publish the trace size of the next iteration, reset cursors.
// tail_n != 0: when every camera sample has been issued and at most tail_n paths are left, hand them to k_tail (launched
// right after raygen in every late iteration; the decision is taken HERE, on the device, so the hand-over does not wait
// for the host's lagged view of the wavefront size).
static __global__ void k_plan(Control* ctl, int nxt, uint32_t* host_n_in, uint32_t tail_n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t survivors = ctl->out_count[nxt];
    const uint64_t remaining = ctl->total_samples - ctl->next_sample;
    uint64_t room = (uint64_t)ctl->pool - survivors;
    const uint32_t n_new = (uint32_t)(remaining < room ? remaining : room);
    ctl->gen_base = survivors;
    ctl->gen_count = n_new;
    ctl->gen_first = ctl->next_sample;
    ctl->next_sample += n_new;
    const uint32_t n_in = survivors + n_new;
    ctl->n_in[nxt] = n_in;
    ctl->rays_traced += n_in;
    ctl->out_count[nxt ^ 1] = 0;
    ctl->trace_next = 0;
    ctl->retry_total += ctl->retry_count;
    ctl->retry_count = 0;
    ctl->retry_next = 0;
    for (int q = 0; q < Q_COUNT; ++q) ctl->queue_count[q] = 0;
    ctl->iteration++;
    ctl->tail_go = (tail_n != 0u && ctl->next_sample >= ctl->total_samples && n_in != 0u && n_in <= tail_n) ? 1u : 0u;
    if (host_n_in) *host_n_in = n_in;
}
// after k_tail: the wavefront of side `side` is finished, later iterations find nothing to do
static __global__ void k_tail_finish(Control* ctl, int side) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || !ctl->tail_go) return;
    ctl->n_in[side] = 0;
    ctl->tail_go = 0;
}

// rewinds the trace cursor and the material queues so the SAME wavefront can be traced again (variant timing)
static __global__ void k_trace_rewind(Control* ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->trace_next = 0;
    ctl->retry_count = 0;
    ctl->retry_next = 0;
    for (int q = 0; q < Q_COUNT; ++q) ctl->queue_count[q] = 0;
}

template <typename R>
__device__ __forceinline__ void store_path(PathRec<R>* p, const PathRec<R>& in);
// the producer of a ray (o, d) also writes its 48 B FilterRec (interval of ray_color: (0.001, inf), ray_casting.rs:119)
template <typename R>
__device__ __forceinline__ void store_filter(FilterRec* dst, V3<R> o, V3<R> d, float bsmall, float bmax) {
    const V3<R> inv = {R(1) / d.x, R(1) / d.y, R(1) / d.z};  // adinv, bvh.rs:111
    const FilterRec r = pack_filter(make_filter_ray<R>(o, d, inv, R(0.001), Num<R>::inf(), bsmall, bmax));
    const int4* s = reinterpret_cast<const int4*>(&r);
    int4* q = reinterpret_cast<int4*>(dst);
    q[0] = s[0]; q[1] = s[1]; q[2] = s[2];
}

// Coalesced write of a warp's run of records.  A lane-private 128 B record written with 16 B stores at a
// 128 B lane stride costs 32 L1 transactions per instruction and throttles the LSU (ncu: stall_lg was 53 % of
// k_raygen).  Instead the lanes stage their records in a per-warp shared-memory tile (XOR-swizzled, conflict
// free) and the warp streams the tile out 512 contiguous bytes per instruction.
//   dst   = first record of the run (the run is contiguous: compaction and raygen both guarantee it)
//   rank  = position of this lane's record in the run, or -1 when the lane has none;  count = records in the run
template <typename R>
__device__ __forceinline__ void warp_store_records(int4* tile, PathRec<R>* dst, int rank, int count, const PathRec<R>& rec) {
    constexpr int NW = (int)(sizeof(PathRec<R>) / 16);
    const int lane = threadIdx.x & 31;
    if (rank >= 0) {
        const int4* src = reinterpret_cast<const int4*>(&rec);
#pragma unroll
        for (int w = 0; w < NW; ++w) tile[rank * NW + (w ^ (rank & (NW - 1)))] = src[w];
    }
    __syncwarp();
    int4* out = reinterpret_cast<int4*>(dst);
    const int total = count * NW;
#pragma unroll
    for (int f = lane; f < 32 * NW; f += 32) {
        if (f < total) {
            const int r = f / NW, w = f % NW;
            out[f] = tile[r * NW + (w ^ (r & (NW - 1)))];
        }
    }
    __syncwarp();
}

// direction of a record (the f64 record keeps a second copy next to the throughput: PathRec, common.cuh)
template <typename R>
__device__ __forceinline__ void set_dir(PathRec<R>& p, V3<R> d) {
    p.dx = d.x; p.dy = d.y; p.dz = d.z;
    if constexpr (sizeof(R) == 8) {
        p.dx2 = d.x; p.dy2 = d.y; p.dz2 = d.z;
    }
}

// raygen: persistent grid-stride over the samples the plan handed out.  Sample-major order
// (g = sample * npix + pixel) keeps neighbouring lanes on neighbouring pixels.  Everything that is
// constant for the launch arrives BY VALUE (constant bank, no load latency): for a keyframe-free camera
// that includes the whole basis, evaluated once on the host with the same IEEE operations
// (host_camera_basis), so the per-sample work is the jitter, the lens and one 128 B record.
template <typename R>
struct RaygenParams {
    uint32_t W, rows_local, row_block, row_rank, row_world, is_static, lens, small;  // small: total samples < 2^32
    uint64_t seed;
    R current_time, shutter_length;
    R from[3], psl[3], pdu[3], pdv[3], du[3], dv[3];
    float bsmall, bmax;  // filter bounds of the committed scene (FilterRec is written with the ray)
};
template <typename R, int MINB>
__global__ void __launch_bounds__(SHADE_BLOCK, MINB) k_raygen(const Control* __restrict__ ctl, const DevCamera* __restrict__ camp,
                                                               const __grid_constant__ RaygenParams<R> rp, PathRec<R>* __restrict__ out,
                                                               FilterRec* __restrict__ filt) {
    const uint32_t n = ctl->gen_count;
    if (n == 0) return;
    const uint32_t base = ctl->gen_base;
    const uint64_t first = ctl->gen_first;
    const uint32_t W = rp.W;
    const uint64_t npix = (uint64_t)W * rp.rows_local;
    __shared__ int4 s_tile[SHADE_BLOCK * (sizeof(PathRec<R>) / 16)];
    int4* tile = s_tile + (threadIdx.x >> 5) * 32 * (sizeof(PathRec<R>) / 16);
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += gridDim.x * blockDim.x) {
        const uint32_t k0 = k & ~31u;  // first sample of this warp's run
        const int count = (int)((n - k0) < 32u ? (n - k0) : 32u);
        PathRec<R> p;
        if (k < n) {
        const uint64_t g = first + k;
        uint32_t sample, lp;
        if (rp.small) {  // 32-bit division is ~5x cheaper than the 64-bit one
            sample = (uint32_t)g / (uint32_t)npix;
            lp = (uint32_t)g - sample * (uint32_t)npix;
        } else {
            sample = (uint32_t)(g / npix);
            lp = (uint32_t)(g - (uint64_t)sample * npix);
        }
        const uint32_t lrow = lp / W, i = lp - lrow * W;
        const uint32_t j = local_to_global_row(lrow, rp.row_block, rp.row_rank, rp.row_world);
        const uint32_t pixel = j * W + i;
        Rng<R> rng(rp.seed, pixel, sample, 0);
        V3<R> o, d;
        R tm;
        if (rp.is_static) {
            // ray_casting.rs:82-104 with the basis hoisted (it does not depend on the sample time)
            tm = rp.current_time + rng.range(R(0), rp.shutter_length);
            const R ox = rng.next() - R(0.5);  // sample_square, camera/mod.rs:369-376
            const R oy = rng.next() - R(0.5);
            const V3<R> psl = {rp.psl[0], rp.psl[1], rp.psl[2]}, pdu = {rp.pdu[0], rp.pdu[1], rp.pdu[2]},
                        pdv = {rp.pdv[0], rp.pdv[1], rp.pdv[2]}, from = {rp.from[0], rp.from[1], rp.from[2]};
            const V3<R> ps = vadd(vadd(psl, vmul((R)i + ox, pdu)), vmul((R)j + oy, pdv));  // get_pixel_pos :64-68
            o = from;
            if (rp.lens) {  // defocus_disk_sample :105-110
                const V3<R> p = random_in_unit_disk(rng);
                o = vadd(vadd(from, vmul(p.x, V3<R>{rp.du[0], rp.du[1], rp.du[2]})), vmul(p.y, V3<R>{rp.dv[0], rp.dv[1], rp.dv[2]}));
            }
            d = vsub(ps, o);
        } else {
            camera_sample<R>(*camp, i, j, rng, o, d, tm);
        }
        p.ox = o.x; p.oy = o.y; p.oz = o.z;
        set_dir(p, d);
        p.tm = tm;
        p.tr = R(1); p.tg = R(1); p.tb = R(1);
        p.bounce = 0;
        p.pixel = pixel;
        p.sample = sample;
        p.fb = lp;
        p.pad0 = p.pad1 = 0;
        store_filter<R>(filt + base + k, o, d, rp.bsmall, rp.bmax);
        }
        warp_store_records<R>(tile, out + base + k0, k < n ? (int)(k - k0) : -1, count, p);
    }
}

// trace: persistent threads with warp-level work fetch and lane refill (trace_persistent), stackless
// threaded traversal, 128-bit node / primitive loads, then classification of the finished rays into their material queues with one atomic per warp per queue.
#ifndef CRB_REFILL
#define CRB_REFILL 24
#endif
template <typename R>
struct RenderTraceIO {
    const DevScene<R>& sc;
    const PathRec<R>* paths;  // read only: the closest hit goes to the material queue
    Control* ctl;
    uint2* queues;  // Q_COUNT queues of `pool` entries, then HIT_QUEUES arrays of `pool` HitEntry<R>
    const FilterRec* filt;
    uint32_t n, pool;
    const uint32_t* remap;  // non-null: work item k is path remap[k] (the retry list of the order-free engine)
    uint32_t* cur;
    // Record loads that do not allocate in L1 (ld.global.cg).  The trace kernels used to WRITE the closest hit into the record
    // they had just read, which as a side effect threw the streamed line out of the L1; since the hit travels in the queue the
    // lines stayed and pushed the scene out: teapot (1.4 MB of tree + triangles against ~77 KB of L1 per SM) lost 6 points of
    // L1 hit rate and 37 % on its big wavefronts (ncu, gpurun_out/r02q_teapot_*.csv).  The shared-memory build keeps its
    // scene out of the L1 and keeps the allocating streaming loads (one miss per record instead of three L2 requests).
    bool bypass_l1;
    __device__ __forceinline__ int4 ld_rec16(const void* p) const { return bypass_l1 ? __ldcg(reinterpret_cast<const int4*>(p)) : ld_stream16(p); }
    __device__ __forceinline__ uint32_t count() const { return n; }
    __device__ __forceinline__ uint32_t* cursor() const { return cur; }
    __device__ __forceinline__ uint32_t path_of(uint32_t k) const { return remap ? remap[k] : k; }
    __device__ __forceinline__ uint32_t item(uint32_t k) const { return path_of(k); }  // what the retry list holds
    __device__ __forceinline__ FilterRay filter(uint32_t k, R, R) const {  // written by the ray's producer
        FilterRec r;
        const int4* s = reinterpret_cast<const int4*>(filt + path_of(k));
        int4* d = reinterpret_cast<int4*>(&r);
        d[0] = ld_rec16(s); d[1] = ld_rec16(s + 1); d[2] = ld_rec16(s + 2);  // streamed once: keep the L1 for the scene
        return unpack_filter(r);
    }
    __device__ __forceinline__ void load(uint32_t k, V3<R>& o, V3<R>& d) const {
        const PathRec<R>* p = paths + path_of(k);  // first 48 B of the record: origin + direction
        if constexpr (sizeof(R) == 8) {
            const int4 a = ld_rec16(&p->ox);
            const int4 b = ld_rec16(&p->oz);
            const int4 c = ld_rec16(&p->dy);
            o = {__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x)};
            d = {__hiloint2double(b.w, b.z), __hiloint2double(c.y, c.x), __hiloint2double(c.w, c.z)};
        } else {
            const int4 a = ld_rec16(&p->ox);
            const int4 b = ld_rec16(&p->dx);
            o = {__int_as_float(a.x), __int_as_float(a.y), __int_as_float(a.z)};
            d = {__int_as_float(b.x), __int_as_float(b.y), __int_as_float(b.z)};
        }
    }
    __device__ __forceinline__ R time(uint32_t k) const { return paths[path_of(k)].tm; }  // ray_casting.rs:84
    __device__ __forceinline__ bool commit_needs_ray() const { return false; }
    __device__ __forceinline__ void commit(bool has, uint32_t k, uint32_t ref, R t, V3<R>, V3<R>, R) const {
        int q = -1;
        uint32_t minfo = 0;  // material index (bits 0..23) | needs-uv (bit 31): saves the shaders a dependent load
        uint32_t i = 0;
        if (has) {
            i = path_of(k);
            if (ref == REF_MISS) {
                q = Q_MISS;
            } else {
                const uint32_t kind = ref_kind(ref);
                const PrimMeta* m = kind == CR_PRIM_SPHERE ? sc.meta[0] : (kind == CR_PRIM_TRIANGLE ? sc.meta[1] : sc.meta[2]);
                const PrimMeta pm = m[ref_index(ref)];
                q = (int)Q_LAMBERTIAN + (pm.mat_kind & MATKIND_MASK);
                minfo = (uint32_t)pm.material | ((pm.mat_kind & MATKIND_NEEDS_UV) ? 0x80000000u : 0u);
            }
        }
        const uint32_t pos = warp_enqueue(ctl->queue_count, q);
        if (q >= 0) __stcs(&queues[(size_t)q * pool + pos], make_uint2(i, minfo));
        if (q >= (int)Q_LAMBERTIAN && q < (int)Q_LAMBERTIAN + HIT_QUEUES) {  // the scatter shaders' (ref, t)
            HitEntry<R>* hits = reinterpret_cast<HitEntry<R>*>(queues + (size_t)Q_COUNT * pool);
            HitEntry<R> he;
            he.t = t;
            he.ref = ref;
            if constexpr (sizeof(R) == 8) {
                he.pad = 0;
                __stcs(reinterpret_cast<int4*>(hits + (size_t)(q - (int)Q_LAMBERTIAN) * pool + pos), *reinterpret_cast<const int4*>(&he));
            } else {
                __stcs(reinterpret_cast<int2*>(hits + (size_t)(q - (int)Q_LAMBERTIAN) * pool + pos), *reinterpret_cast<const int2*>(&he));
            }
        }
    }
};

// remap == nullptr: the whole wavefront of side `side`; otherwise the retry list the order-free kernel left behind
template <typename R, int REFILL, int MINB, bool ANIM>
__global__ void __launch_bounds__(TRACE_BLOCK, MINB) k_trace(DevScene<R> sc, const PathRec<R>* __restrict__ paths, Control* __restrict__ ctl,
                                                        int side, uint2* __restrict__ queues, const FilterRec* __restrict__ filt, uint32_t pool,
                                                        const uint32_t* __restrict__ remap) {
    __shared__ LaneSlots<R, TRACE_BLOCK, ANIM> slots;
    RenderTraceIO<R> io{sc, paths, ctl, queues, filt, remap ? ctl->retry_count : ctl->n_in[side], pool, remap,
                        remap ? &ctl->retry_next : &ctl->trace_next, sc.rec_bypass_l1 != 0};
    if (io.n == 0u) return;
    trace_persistent<R, REFILL, TRACE_BLOCK, ANIM>(sc, R(0.001), Num<R>::inf(), io, &slots);  // ray_casting.rs:119
}

// The order-free engine over the wavefront (fast_trace.cuh); undecidable rays go to retry_list for k_trace above.
template <typename R, int MINB>
__global__ void __launch_bounds__(TRACE_BLOCK, MINB) k_trace_fast(DevScene<R> sc, const PathRec<R>* __restrict__ paths, Control* __restrict__ ctl,
                                                             int side, uint2* __restrict__ queues, const FilterRec* __restrict__ filt,
                                                             uint32_t pool, uint32_t* __restrict__ retry_list) {
    __shared__ FastSlots<R, TRACE_BLOCK> slots;
    RenderTraceIO<R> io{sc, paths, ctl, queues, filt, ctl->n_in[side], pool, nullptr, &ctl->trace_next, sc.rec_bypass_l1 != 0};
    fast_trace_persistent<R, TRACE_BLOCK, false>(sc, R(0.001), Num<R>::inf(), io, &slots, retry_list, &ctl->retry_count);
}

// Small scenes: the same engine with the search tree in shared memory (SmemTree, fast_trace.cuh).  One CTA of 28 warps
// per SM (the 72-register budget of the default build) so that the SM holds one copy of the tree.
static constexpr int FAST_BIG_BLOCK = 896;
static constexpr int FAST_BIG_LEVELS = 6;  // stack levels per lane in the lane table (deeper: local memory)
static constexpr size_t FAST_SMEM_LIMIT = 227u * 1024u;
struct FastSmemLayout {
    uint32_t nodes, prims, spheres32, spheres, tris, quads, leafbox, total;
};
template <typename R>
static __host__ __device__ FastSmemLayout fast_smem_layout(uint32_t n_fast_nodes, uint32_t n_fast_prims, uint32_t n_sph, uint32_t n_tri, uint32_t n_quad,
                                                           uint32_t node_stride = 64u) {
    auto up = [](uint32_t b) { return (b + 127u) & ~127u; };
    FastSmemLayout l;
    l.nodes = up((uint32_t)sizeof(FastSlots<R, FAST_BIG_BLOCK, FAST_BIG_LEVELS>));
    l.prims = l.nodes + up(n_fast_nodes * node_stride);
    l.spheres32 = l.prims + up(n_fast_prims * 8u);
    l.spheres = l.spheres32 + up(n_sph * 16u);
    l.tris = l.spheres + up(n_sph * (uint32_t)sizeof(SphereRec<R>));
    l.quads = l.tris + up(n_tri * (uint32_t)sizeof(TriRec<R>));
    l.leafbox = l.quads + up(n_quad * ((uint32_t)sizeof(QuadRec<R>) + 16u));  // quad records padded by 16 B (SmemTree::quad_stride)
    l.total = l.leafbox + up(n_fast_prims * (uint32_t)sizeof(LeafBox<R>));
    return l;
}
template <typename R>
__global__ void __launch_bounds__(FAST_BIG_BLOCK, 1) k_trace_fast_smem(DevScene<R> sc, const PathRec<R>* __restrict__ paths, Control* __restrict__ ctl,
                                                                       int side, uint2* __restrict__ queues, const FilterRec* __restrict__ filt,
                                                                       uint32_t pool, uint32_t* __restrict__ retry_list, uint32_t n_fast_nodes,
                                                                       uint32_t n_fast_prims, uint32_t n_sph, uint32_t n_tri, uint32_t n_quad,
                                                                       uint32_t node_stride) {
    extern __shared__ __align__(128) unsigned char fast_smem[];
    if (ctl->n_in[side] == 0u) return;  // nothing to trace: skip the copy of the tree
    typedef FastSlots<R, FAST_BIG_BLOCK, FAST_BIG_LEVELS> Slots;
    Slots* slots = reinterpret_cast<Slots*>(fast_smem);
    const FastSmemLayout l = fast_smem_layout<R>(n_fast_nodes, n_fast_prims, n_sph, n_tri, n_quad, node_stride);
    {
        auto copy16 = [&](uint32_t dst_off, const void* src, uint32_t n16) {
            int4* d = reinterpret_cast<int4*>(fast_smem + dst_off);
            const int4* s = reinterpret_cast<const int4*>(src);
            for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) d[i] = __ldg(s + i);
        };
        {  // 64 B node records at a stride of node_stride bytes (80: see SmemTree::node_stride)
            int4* d = reinterpret_cast<int4*>(fast_smem + l.nodes);
            const int4* s = reinterpret_cast<const int4*>(sc.fast_nodes);
            const uint32_t per = node_stride >> 4;
            for (uint32_t i = threadIdx.x; i < n_fast_nodes * 4u; i += blockDim.x) d[(i >> 2) * per + (i & 3u)] = __ldg(s + i);
        }
        uint2* dp = reinterpret_cast<uint2*>(fast_smem + l.prims);
        LeafBox<R>* db = reinterpret_cast<LeafBox<R>*>(fast_smem + l.leafbox);
        for (uint32_t i = threadIdx.x; i < n_fast_prims; i += blockDim.x) {
            const uint2 e = __ldg(sc.fast_prims + i);
            dp[i] = e;
            const NodeRec<R>* n = sc.nodes + (e.y >> 1);  // the primitive's reference leaf node: its box confirms candidates
            LeafBox<R> b;
            b.xmin = n->xmin; b.xmax = n->xmax; b.ymin = n->ymin; b.ymax = n->ymax; b.zmin = n->zmin; b.zmax = n->zmax;
            db[i] = b;
        }
        copy16(l.spheres32, sc.spheres32, n_sph);
        copy16(l.spheres, sc.spheres, n_sph * (uint32_t)(sizeof(SphereRec<R>) / 16));
        copy16(l.tris, sc.tris, n_tri * (uint32_t)(sizeof(TriRec<R>) / 16));
        {  // quad records at a stride of sizeof + 16 B (see SmemTree::quad_stride)
            int4* d = reinterpret_cast<int4*>(fast_smem + l.quads);
            const int4* s = reinterpret_cast<const int4*>(sc.quads);
            constexpr uint32_t W = (uint32_t)(sizeof(QuadRec<R>) / 16);
            for (uint32_t i = threadIdx.x; i < n_quad * W; i += blockDim.x) d[(i / W) * (W + 1u) + (i % W)] = __ldg(s + i);
        }
    }
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(fast_smem);
    RenderTraceIO<R> io{sc, paths, ctl, queues, filt, ctl->n_in[side], pool, nullptr, &ctl->trace_next, false};
    SmemTree tree;
    tree.node_stride = node_stride;
    tree.quad_stride = (uint32_t)sizeof(QuadRec<R>) + 16u;
    tree.nodes = base + l.nodes; tree.prims = base + l.prims; tree.spheres32 = base + l.spheres32; tree.spheres = base + l.spheres;
    tree.tris = base + l.tris; tree.quads = base + l.quads; tree.leafbox = base + l.leafbox;
    fast_trace_persistent<R, FAST_BIG_BLOCK, true, FAST_BIG_LEVELS>(sc, R(0.001), Num<R>::inf(), io, slots, retry_list, &ctl->retry_count, tree);
}

// fixed-point accumulation: order independent => bit-reproducible for any schedule and GPU count
static __device__ __forceinline__ void fb_add(unsigned long long* fb, uint32_t idx, double r, double g, double b, double scale) {
    atomicAdd(fb + 3ull * idx + 0, __double2ull_rn(r * scale));
    atomicAdd(fb + 3ull * idx + 1, __double2ull_rn(g * scale));
    atomicAdd(fb + 3ull * idx + 2, __double2ull_rn(b * scale));
}

template <typename R>
__device__ __forceinline__ void load_path(const PathRec<R>* p, PathRec<R>& out) {
    const int4* s = reinterpret_cast<const int4*>(p);
    int4* d = reinterpret_cast<int4*>(&out);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PathRec<R>) / 16); ++i) d[i] = s[i];
}
template <typename R>
__device__ __forceinline__ void store_path(PathRec<R>* p, const PathRec<R>& in) {
    int4* d = reinterpret_cast<int4*>(p);
    const int4* s = reinterpret_cast<const int4*>(&in);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PathRec<R>) / 16); ++i) d[i] = s[i];
}

// miss: ray_color's skybox arm; the path ends and thr * sky is accumulated
template <typename R>
__global__ void __launch_bounds__(SHADE_BLOCK, 8) k_shade_miss(DevScene<R> sc, const PathRec<R>* __restrict__ in,
                                                             const Control* __restrict__ ctl, const uint2* __restrict__ queue,
                                                             unsigned long long* __restrict__ fb, double fb_scale) {
    const uint32_t n = ctl->queue_count[Q_MISS];
    const bool cl = sc.clamp_colors != 0;
    // Two paths per thread and trip, all loads of both issued before anything is used: the kernel is a chain of dependent
    // long-latency accesses (queue entry -> record -> three framebuffer reductions, 58 % of which miss the L2) and ran at 16 %
    // issue, 62 % DRAM, 47 % L2 — bound by none of them (ncu, profiles/shade_book1_r02g.md), i.e. by memory-level parallelism.
    constexpr int U = 2;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x; k0 < n; k0 += U * stride) {
        uint32_t idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t k = k0 + (uint32_t)u * stride;
            idx[u] = k < n ? __ldcs(&queue[k].x) : 0xFFFFFFFFu;
        }
        int4 w[U][sizeof(R) == 8 ? 4 : 2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (idx[u] == 0xFFFFFFFFu) continue;
            const PathRec<R>* rec = in + idx[u];
            if constexpr (sizeof(R) == 8) {  // the second 64 B half only: throughput, the direction's copy, pixel / sample / bounce / fb
                const int4* q = reinterpret_cast<const int4*>(&rec->tr);
                w[u][0] = __ldcs(q); w[u][1] = __ldcs(q + 1); w[u][2] = __ldcs(q + 2); w[u][3] = __ldcs(q + 3);
            } else {
                w[u][0] = __ldcs(reinterpret_cast<const int4*>(&rec->dx));
                w[u][1] = __ldcs(reinterpret_cast<const int4*>(&rec->tr));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (idx[u] == 0xFFFFFFFFu) continue;
            V3<R> d, thr;
            uint32_t fbi;
            if constexpr (sizeof(R) == 8) {
                thr = {__hiloint2double(w[u][0].y, w[u][0].x), __hiloint2double(w[u][0].w, w[u][0].z), __hiloint2double(w[u][1].y, w[u][1].x)};
                d = {__hiloint2double(w[u][1].w, w[u][1].z), __hiloint2double(w[u][2].y, w[u][2].x), __hiloint2double(w[u][2].w, w[u][2].z)};
                fbi = (uint32_t)w[u][3].w;
            } else {
                d = {__int_as_float(w[u][0].x), __int_as_float(w[u][0].y), __int_as_float(w[u][0].z)};
                thr = {__int_as_float(w[u][1].x), __int_as_float(w[u][1].y), __int_as_float(w[u][1].z)};
                fbi = (uint32_t)w[u][1].w;
            }
            const V3<R> sky = sky_color<R>(sc, d);
            const V3<R> c = col_mul(thr, sky, cl);
            fb_add(fb, fbi, (double)c.x, (double)c.y, (double)c.z, fb_scale);
        }
    }
}

// EXTENSION emissive: the path ends with thr * emit
template <typename R>
__global__ void __launch_bounds__(SHADE_BLOCK) k_shade_emissive(DevScene<R> sc, const PathRec<R>* __restrict__ in,
                                                                 const Control* __restrict__ ctl,
                                                                 const uint2* __restrict__ queue,
                                                                 unsigned long long* __restrict__ fb, double fb_scale) {
    const uint32_t n = ctl->queue_count[Q_EMISSIVE];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint2 e = queue[k];
        const PathRec<R>* rec = in + e.x;
        const R tr = rec->tr, tg = rec->tg, tb = rec->tb;
        const uint32_t fbi = rec->fb;
        const DevMaterial& mat = sc.mats[e.y & 0x00FFFFFFu];
        fb_add(fb, fbi, (double)(tr * (R)mat.emit[0]), (double)(tg * (R)mat.emit[1]), (double)(tb * (R)mat.emit[2]), fb_scale);
    }
}

// scatter kernels: one per material so a warp runs one BSDF.  MAT selects the arm at compile time.
// One scatter event (Materials::scatter, src/materials/mod.rs:23-29) on a path whose closest hit is
// (p.ref, p.t): rebuilds the HitRecord, draws from the (pixel, sample, bounce) stream, and on survival
// turns p into the scattered ray.  Shared by the material-sorted shade kernels and the tail kernel.
template <typename R, int MAT, bool ANIM>
__device__ __forceinline__ bool scatter_path(const DevScene<R>& sc, PathRec<R>& p, uint32_t hit_ref, R hit_t, uint32_t minfo, uint64_t seed,
                                             uint32_t max_depth) {
    const bool cl = sc.clamp_colors != 0;
    const V3<R> o = {p.ox, p.oy, p.oz}, d = {p.dx, p.dy, p.dz};
    // u, v are needed only when a Lambertian's texture tree reaches an image
    const HitInfo<R> h = finalize_geom<R, ANIM>(sc, hit_ref, hit_t, o, d, (minfo >> 31) != 0u, p.tm);
    const DevMaterial& mat = sc.mats[minfo & 0x00FFFFFFu];
    const uint32_t bounce = p.bounce + 1;  // this is the bounce-th hit of the path
    Rng<R> g(seed, p.pixel, p.sample, bounce);
    V3<R> att, nd;
    bool alive;
    if (MAT == CR_MAT_LAMBERTIAN) {  // lambertian.rs:40-61
        V3<R> dir = vadd(h.n, random_unit_vector(g));
        if (vnear_zero(dir)) dir = h.n;
        nd = dir;
        att = col_div(tex_value<R>(sc, mat.tex, h.u, h.v, h.p), (R)mat.scatter_prob, true);
        alive = g.next() <= (R)mat.scatter_prob;
    } else if (MAT == CR_MAT_METAL) {  // metal.rs:29-42
        const V3<R> refl = vreflect(d, h.n);
        nd = vadd(vunit(refl), vmul((R)mat.fuzz, random_unit_vector(g)));
        att = {(R)mat.albedo[0], (R)mat.albedo[1], (R)mat.albedo[2]};
        alive = vdot(nd, h.n) > R(0);
    } else {  // dielectric.rs:30-55
        att = {R(1), R(1), R(1)};
        const R ri = h.front ? R(1) / (R)mat.ior : (R)mat.ior;
        const V3<R> ud = vunit(d);
        const R cos_theta = -(Num<R>::min_(vdot(ud, h.n), R(1)));
        const R sin_theta = Num<R>::sqrt_(R(1) - cos_theta * cos_theta);
        bool refl = ri * sin_theta > R(1);
        if (!refl) {
            R r0 = (R(1) - ri) / (R(1) + ri);
            r0 = r0 * r0;
            const R x = R(1) - cos_theta;
            const R x2 = x * x;
            const R x5 = x * (x2 * x2);
            refl = (r0 + (R(1) - r0) * x5) > g.next();
        }
        nd = refl ? vreflect(ud, h.n) : vrefract(ud, h.n, ri);
        alive = true;
    }
    // ray_color(depth == 0) returns black: a path that has used max_depth hits contributes nothing
    if (bounce >= max_depth) alive = false;
    if (alive) {
        const V3<R> thr = col_mul(V3<R>{p.tr, p.tg, p.tb}, att, cl);
        p.ox = h.p.x; p.oy = h.p.y; p.oz = h.p.z;
        set_dir(p, nd);
        p.tr = thr.x; p.tg = thr.y; p.tb = thr.z;
        p.bounce = bounce;
    }
    return alive;
}

// scatter kernels: one per material so a warp runs one BSDF; survivors are compacted into the other side
template <typename R, int MAT, int MINB, bool ANIM>
__global__ void __launch_bounds__(SHADE_BLOCK, MINB) k_shade_scatter(DevScene<R> sc, const PathRec<R>* __restrict__ in,
                                                                      PathRec<R>* __restrict__ out, Control* __restrict__ ctl, int nxt,
                                                                      const uint2* __restrict__ queue, const HitEntry<R>* __restrict__ hits,
                                                                      FilterRec* __restrict__ filt_out, uint64_t seed, uint32_t max_depth) {
    const uint32_t n = ctl->queue_count[Q_LAMBERTIAN + MAT];
    if (n == 0) return;
    __shared__ int4 s_tile[SHADE_BLOCK * (sizeof(PathRec<R>) / 16)];
    int4* tile = s_tile + (threadIdx.x >> 5) * 32 * (sizeof(PathRec<R>) / 16);
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += gridDim.x * blockDim.x) {
        bool alive = false;
        PathRec<R> p;
        if (k < n) {
            const uint2 e = queue[k];
            const HitEntry<R> he = hits[k];
            load_path(in + e.x, p);
            alive = scatter_path<R, MAT, ANIM>(sc, p, he.ref, he.t, e.y, seed, max_depth);
        }
        const uint32_t amask = __ballot_sync(0xffffffffu, alive);
        if (amask != 0u) {  // warp-uniform
            const uint32_t pos = warp_append(&ctl->out_count[nxt], alive);
            const int rank = (int)__popc(amask & ((1u << (threadIdx.x & 31)) - 1u));
            const uint32_t start = __shfl_sync(0xffffffffu, pos - (uint32_t)rank, __ffs(amask) - 1);
            if (alive) store_filter<R>(filt_out + pos, V3<R>{p.ox, p.oy, p.oz}, V3<R>{p.dx, p.dy, p.dz}, sc.bsmall, sc.bmax);
            warp_store_records<R>(tile, out + start, alive ? rank : -1, (int)__popc(amask), p);
        }
    }
}

// tail: when only a few thousand paths are left (and no camera samples), another ~25-45 wavefront iterations
// of 8 launches each would be pure launch overhead.  One launch finishes them: a WARP owns 32 paths and follows
// them to their end, alternating a warp-cooperative closest-hit (trace_warp_batch: the engine's phases) with the
// scatter of every lane's own material, using the same device routines, so every path is unchanged.
// (The first version gave each THREAD one path and its own traversal state machine: lanes in different states
// serialised each other, 1.6 ms per frame at 11 % warps active, profiles/misc_r01c.md.)
// FAST: the trace step is the order-free engine's (static scenes with a search tree); a lane whose ray it hands back
// is traced by the reference-order walk in the same bounce.
template <typename R, bool ANIM, bool FAST>
__global__ void __launch_bounds__(TRACE_BLOCK, 4) k_tail(DevScene<R> sc, const PathRec<R>* __restrict__ paths, Control* __restrict__ ctl,
                                                         int side, uint64_t seed, uint32_t max_depth,
                                                         unsigned long long* __restrict__ fb, double fb_scale) {
    __shared__ LaneSlots<R, TRACE_BLOCK, ANIM> slots;
    __shared__ FastSlots<R, FAST ? TRACE_BLOCK : 1> fslots;
    if (!ctl->tail_go) return;  // k_plan decides (device side) when the tail takes over
    const uint32_t n = ctl->n_in[side];
    const bool cl = sc.clamp_colors != 0;
    unsigned long long traced = 0;  // segments beyond each path's first (k_plan already counted that one)
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += gridDim.x * blockDim.x) {
        PathRec<R> p;
        bool alive = k < n;
        if (alive) load_path(paths + k, p);
        for (bool first = true; __any_sync(0xffffffffu, alive); first = false) {
            const V3<R> o = {p.ox, p.oy, p.oz}, d = {p.dx, p.dy, p.dz};
            uint32_t ref;
            R t;
            if constexpr (FAST) {
                bool retry;
                fast_trace_warp_batch<R, TRACE_BLOCK>(sc, R(0.001), Num<R>::inf(), alive, o, d, &fslots, ref, t, retry);
                if (__any_sync(0xffffffffu, retry)) {
                    uint32_t ref2;
                    R t2;
                    trace_warp_batch<R, TRACE_BLOCK, ANIM>(sc, R(0.001), Num<R>::inf(), retry, o, d, p.tm, &slots, ref2, t2);
                    if (retry) {
                        ref = ref2;
                        t = t2;
                    }
                }
            } else {
                trace_warp_batch<R, TRACE_BLOCK, ANIM>(sc, R(0.001), Num<R>::inf(), alive, o, d, p.tm, &slots, ref, t);  // ray_casting.rs:119
            }
            if (!alive) continue;
            if (!first) ++traced;
            if (ref == REF_MISS) {  // ray_casting.rs:133-151
                const V3<R> c = col_mul(V3<R>{p.tr, p.tg, p.tb}, sky_color<R>(sc, d), cl);
                fb_add(fb, p.fb, (double)c.x, (double)c.y, (double)c.z, fb_scale);
                alive = false;
                continue;
            }
            const uint32_t kind = ref_kind(ref);
            const PrimMeta pm = (kind == CR_PRIM_SPHERE ? sc.meta[0] : (kind == CR_PRIM_TRIANGLE ? sc.meta[1] : sc.meta[2]))[ref_index(ref)];
            const uint32_t minfo = (uint32_t)pm.material | ((pm.mat_kind & MATKIND_NEEDS_UV) ? 0x80000000u : 0u);
            const int mk = pm.mat_kind & MATKIND_MASK;
            if (mk == CR_MAT_LAMBERTIAN) alive = scatter_path<R, CR_MAT_LAMBERTIAN, ANIM>(sc, p, ref, t, minfo, seed, max_depth);
            else if (mk == CR_MAT_METAL) alive = scatter_path<R, CR_MAT_METAL, ANIM>(sc, p, ref, t, minfo, seed, max_depth);
            else if (mk == CR_MAT_DIELECTRIC) alive = scatter_path<R, CR_MAT_DIELECTRIC, ANIM>(sc, p, ref, t, minfo, seed, max_depth);
            else {  // EXTENSION emissive
                const DevMaterial& mat = sc.mats[pm.material];
                fb_add(fb, p.fb, (double)(p.tr * (R)mat.emit[0]), (double)(p.tg * (R)mat.emit[1]), (double)(p.tb * (R)mat.emit[2]), fb_scale);
                alive = false;
            }
        }
    }
    // ray segments traced here (CrStats.rays counts world.hit calls)
    for (int off = 16; off > 0; off >>= 1) traced += __shfl_down_sync(0xffffffffu, traced, off);
    if ((threadIdx.x & 31) == 0 && traced) atomicAdd(reinterpret_cast<unsigned long long*>(&ctl->rays_traced), traced);
}

// resolve: average_samples (ray_casting.rs:154-173) + Display for Color (utils.rs:422-438)
static __global__ void k_resolve(const unsigned long long* __restrict__ fb, uint32_t npix_local, uint32_t W, uint32_t row_block,
                          uint32_t row_rank, uint32_t row_world, int packed, double inv_scale, double spp, int clamp_out,
                          double* __restrict__ out_rgb, uint8_t* __restrict__ out_rgb8) {
    for (uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x; lp < npix_local; lp += gridDim.x * blockDim.x) {
        const uint32_t lrow = lp / W, i = lp % W;
        const uint32_t j = packed ? lrow : local_to_global_row(lrow, row_block, row_rank, row_world);
        const size_t o = ((size_t)j * W + i) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = (double)fb[3ull * lp + c] * inv_scale;
            v = v / spp;
            if (clamp_out) v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            if (out_rgb) out_rgb[o + c] = v;
            if (out_rgb8) {
                const double b = 255.0 * sqrt(v);
                out_rgb8[o + c] = (uint8_t)(b >= 255.0 ? 255u : (uint32_t)b);
            }
        }
    }
}

// trace_batch: Hittables::hit on caller-supplied rays, full HitRecord out.  Same traversal engine as
// the render path.
template <typename R, bool ANIM>
struct BatchTraceIO {
    const DevScene<R>& sc;
    const double* rays;
    CrHit* out;
    uint32_t* cur;
    uint32_t n;
    const uint32_t* remap;  // non-null: work item k is ray remap[k] (retry list of the order-free engine)
    __device__ __forceinline__ uint32_t count() const { return n; }
    __device__ __forceinline__ uint32_t* cursor() const { return cur; }
    __device__ __forceinline__ uint32_t ray_of(uint32_t k) const { return remap ? remap[k] : k; }
    __device__ __forceinline__ uint32_t item(uint32_t k) const { return ray_of(k); }
    __device__ __forceinline__ FilterRay filter(uint32_t k, R tmin, R tmax) const {
        V3<R> o, d;
        load(k, o, d);
        const V3<R> inv = {R(1) / d.x, R(1) / d.y, R(1) / d.z};  // adinv, bvh.rs:111
        return make_filter_ray<R>(o, d, inv, tmin, tmax, sc.bsmall, sc.bmax);
    }
    __device__ __forceinline__ void load(uint32_t k, V3<R>& o, V3<R>& d) const {
        const double* r = rays + 7ull * ray_of(k);
        o = {(R)r[0], (R)r[1], (R)r[2]};
        d = {(R)r[3], (R)r[4], (R)r[5]};
    }
    __device__ __forceinline__ R time(uint32_t k) const { return (R)rays[7ull * ray_of(k) + 6]; }
    __device__ __forceinline__ bool commit_needs_ray() const { return true; }
    __device__ __forceinline__ void commit(bool has, uint32_t k, uint32_t ref, R t, V3<R> o, V3<R> d, R tm) const {
        if (!has) return;
        CrHit h;
        if (ref == REF_MISS) {
            h.prim_index = -1; h.obj_id = -1; h.front_face = 0; h.material = -1;
            h.t = 0.0; h.p[0] = h.p[1] = h.p[2] = 0.0; h.n[0] = h.n[1] = h.n[2] = 0.0; h.u = h.v = 0.0;
        } else {
            const HitInfo<R> hi = finalize_hit<R, ANIM>(sc, ref, t, o, d, tm);
            h.prim_index = hi.prim_index; h.obj_id = hi.obj_id; h.front_face = hi.front ? 1 : 0; h.material = hi.material;
            h.t = (double)t;
            h.p[0] = (double)hi.p.x; h.p[1] = (double)hi.p.y; h.p[2] = (double)hi.p.z;
            h.n[0] = (double)hi.n.x; h.n[1] = (double)hi.n.y; h.n[2] = (double)hi.n.z;
            h.u = (double)hi.u; h.v = (double)hi.v;
        }
        out[ray_of(k)] = h;
    }
};
// cursor[0] = work cursor, cursor[1] = retry count (written by the order-free kernel), cursor[2] = retry work cursor
template <typename R, bool ANIM>
__global__ void __launch_bounds__(TRACE_BLOCK) k_trace_batch(DevScene<R> sc, const double* __restrict__ rays, uint32_t n, double tmin,
                                                              double tmax, CrHit* __restrict__ out, uint32_t* __restrict__ cursor,
                                                              const uint32_t* __restrict__ remap) {
    __shared__ LaneSlots<R, TRACE_BLOCK, ANIM> slots;
    BatchTraceIO<R, ANIM> io{sc, rays, out, remap ? cursor + 2 : cursor, remap ? cursor[1] : n, remap};
    if (io.n == 0u) return;
    trace_persistent<R, CRB_REFILL, TRACE_BLOCK, ANIM>(sc, (R)tmin, (R)tmax, io, &slots);
}
template <typename R>
__global__ void __launch_bounds__(TRACE_BLOCK) k_trace_batch_fast(DevScene<R> sc, const double* __restrict__ rays, uint32_t n, double tmin,
                                                                   double tmax, CrHit* __restrict__ out, uint32_t* __restrict__ cursor,
                                                                   uint32_t* __restrict__ retry_list) {
    __shared__ FastSlots<R, TRACE_BLOCK> slots;
    BatchTraceIO<R, false> io{sc, rays, out, cursor, n, nullptr};
    fast_trace_persistent<R, TRACE_BLOCK, false>(sc, (R)tmin, (R)tmax, io, &slots, retry_list, cursor + 1);
}

// ---- host side: typed view of the scene + wavefront driver -----------------------------------------
template <typename R>
static DevScene<R> make_dev_scene(const SceneDeviceData& s) {
    DevScene<R> d;
    const int k = (sizeof(R) == 8) ? 0 : 1;
    d.nodes = reinterpret_cast<const NodeRec<R>*>(s.nodes[k]);
    d.nodes32 = reinterpret_cast<const NodeRec<float>*>(s.nodes[1]);
    d.spheres32 = reinterpret_cast<const SphereRec<float>*>(s.spheres[1]);
    d.bmax = s.bmax;
    d.bsmall = s.bsmall;
    d.spheres = reinterpret_cast<const SphereRec<R>*>(s.spheres[k]);
    d.tris = reinterpret_cast<const TriRec<R>*>(s.tris[k]);
    d.quads = reinterpret_cast<const QuadRec<R>*>(s.quads[k]);
    for (int i = 0; i < 3; ++i) d.meta[i] = s.meta[i];
    d.mats = s.mats;
    d.anim_keys = s.anim_keys;
    d.sphere_track = s.sphere_track;
    d.tri_track = s.tri_track;
    d.tri_anim_slot = s.tri_anim_slot;
    d.tri_anim_verts = s.tri_anim_verts;
    d.texs = s.texs;
    d.images = s.images;
    d.fast_nodes = reinterpret_cast<const FastNodeRec*>(s.fast_nodes);
    d.fast_prims = reinterpret_cast<const uint2*>(s.fast_prims);
    d.fast_margin_k = 9.5367431640625e-07f;  // 2^-20 (fast_trace.cuh)
    if (const char* e = getenv("CRB_FAST_MARGIN")) d.fast_margin_k = (float)atof(e);
    d.refill = CRB_REFILL;
    if (const char* e = getenv("CRB_REFILL_RT")) d.refill = atoi(e) > 0 && atoi(e) <= 32 ? atoi(e) : CRB_REFILL;
    d.n_nodes = s.n_nodes;
    d.sky_kind = s.sky_kind;
    d.sky_image = s.sky_image;
    d.clamp_colors = s.clamp_colors;
    d.node_slice = s.node_slice;
    if (const char* e = getenv("CRB_NODE_SLICE")) d.node_slice = atoi(e) > 0 ? atoi(e) : 8;
    // (a scene far beyond the caches gains nothing from a protected L1: 10 M triangles, +0.8 % with the bypass, A/B r02r)
    d.rec_bypass_l1 = ((uint64_t)s.n_prims[0] + s.n_prims[1] + s.n_prims[2]) <= (1u << 18) ? 1 : 0;
    if (const char* e = getenv("CRB_REC_BYPASS")) d.rec_bypass_l1 = atoi(e) != 0;
    d.min_node_lanes = 8;
    if (const char* e = getenv("CRB_MIN_LANES")) d.min_node_lanes = atoi(e);
    d.free_pass_nodes = 31;
    if (const char* e = getenv("CRB_FREE_PASS")) d.free_pass_nodes = (uint32_t)strtoul(e, nullptr, 10);
    d.free_pass_k = 4.f;
    if (const char* e = getenv("CRB_FREE_PASS_K")) d.free_pass_k = (float)atof(e);
    d.strict_boxes = s.strict_boxes;
    if (s.strict_boxes) {
        d.free_pass_nodes = 0u;
        d.free_pass_k = std::numeric_limits<float>::infinity();
    }
    return d;
}

#define CRB_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return CR_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// The order-free engine serves static scenes that have a search tree; object keyframes (primitives leave their
// construction-time boxes, which only the reference tree's semantics cover) and explicit requests keep reference order.
static bool use_fast_engine(const SceneDeviceData& s, int reference_order) {
    if (const char* e = getenv("CRB_TRAVERSAL")) {
        if (e[0] == 'r') return false;
    }
    return !reference_order && s.fast_nodes != nullptr && s.anim_keys == nullptr && s.n_nodes != 0u;
}

template <typename F>
static int persistent_grid(F kernel, int block, int num_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * num_sms;  // a whole number of resident CTAs per SM (148 SMs on B200)
}

// d_cursor: 4 words (work cursor, retry count, retry cursor, spare); d_retry: n words (order-free engine only)
template <typename R>
int trace_batch_impl(const SceneDeviceData& s, const double* d_rays, size_t n, double tmin, double tmax, CrHit* d_out,
                     uint32_t* d_cursor, uint32_t* d_retry, int reference_order, uint32_t* h_retried, cudaStream_t stream, std::string& err) {
    if (h_retried) *h_retried = 0;
    if (n == 0) return CR_OK;
    if (n > 0xFFFFFF00ull) {
        err = "trace_batch: more than 2^32 rays in one call";
        return CR_ERR_LIMIT;
    }
    const DevScene<R> sc = make_dev_scene<R>(s);
    const bool animated = s.anim_keys != nullptr;
    auto kern = animated ? k_trace_batch<R, true> : k_trace_batch<R, false>;
    int grid = persistent_grid(kern, TRACE_BLOCK, s.num_sms);
    const size_t need = (n + TRACE_BLOCK - 1) / TRACE_BLOCK;
    if ((size_t)grid > need) grid = (int)need;
    CRB_CUDA(cudaMemsetAsync(d_cursor, 0, 4 * sizeof(uint32_t), stream));
    const bool fast = use_fast_engine(s, reference_order) && d_retry != nullptr;
    if (fast) {
        int gf = persistent_grid(k_trace_batch_fast<R>, TRACE_BLOCK, s.num_sms);
        if ((size_t)gf > need) gf = (int)need;
        k_trace_batch_fast<R><<<gf, TRACE_BLOCK, 0, stream>>>(sc, d_rays, (uint32_t)n, tmin, tmax, d_out, d_cursor, d_retry);
        // the rays it could not decide, in reference order (usually none: the launch then returns at once)
        kern<<<std::min(grid, s.num_sms * 2), TRACE_BLOCK, 0, stream>>>(sc, d_rays, (uint32_t)n, tmin, tmax, d_out, d_cursor, d_retry);
        if (h_retried) CRB_CUDA(cudaMemcpyAsync(h_retried, d_cursor + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    } else {
        kern<<<grid, TRACE_BLOCK, 0, stream>>>(sc, d_rays, (uint32_t)n, tmin, tmax, d_out, d_cursor, nullptr);
    }
    CRB_CUDA(cudaGetLastError());
    return CR_OK;
}

// Host evaluation of the launch-constant camera terms with the reference's operations (the library's
// host code is compiled with -ffp-contract=off; IEEE add/mul/div/sqrt give the device's bits).
//   current_time, shutter_length: ray_casting.rs:77-79;  basis: rendering_compute.rs:5-111
template <typename R>
static void host_camera_constants(const CrCamera& c, RaygenParams<R>& rp) {
    struct H3 { R x, y, z; };
    auto neg = [](H3 a) { return H3{-a.x, -a.y, -a.z}; };
    auto add = [](H3 a, H3 b) { return H3{a.x + b.x, a.y + b.y, a.z + b.z}; };
    auto sub = [&](H3 a, H3 b) { return add(a, neg(b)); };
    auto mul = [](R s, H3 v) { return H3{s * v.x, s * v.y, s * v.z}; };
    auto divs = [&](H3 v, R s) { return mul(R(1) / s, v); };
    auto len2 = [](H3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; };
    auto cross = [](H3 v, H3 o) { return H3{v.y * o.z - v.z * o.y, v.z * o.x - v.x * o.z, v.x * o.y - v.y * o.x}; };
    auto unit = [&](H3 v) { return divs(v, (R)std::sqrt(len2(v))); };
    rp.current_time = (R)c.frame * (R(1) / (R)c.frame_rate);
    rp.shutter_length = ((R)c.shutter_angle / R(360)) * (R(1) / (R)c.frame_rate);
    const H3 from = {(R)c.look_from[0], (R)c.look_from[1], (R)c.look_from[2]};
    const H3 at = {(R)c.look_at[0], (R)c.look_at[1], (R)c.look_at[2]};
    const H3 vup = {(R)c.vup[0], (R)c.vup[1], (R)c.vup[2]};
    const H3 w = unit(sub(from, at));
    const H3 u = unit(cross(vup, w));
    const H3 v = cross(w, u);
    const H3 vu = mul((R)c.viewport_width, u);
    const H3 vv = mul((R)c.viewport_height, neg(v));
    const H3 pdu = divs(vu, (R)c.image_width);
    const H3 pdv = divs(vv, (R)c.image_height);
    const H3 ul = sub(sub(sub(from, mul((R)c.focus_dist, w)), divs(vu, R(2))), divs(vv, R(2)));
    const H3 psl = add(ul, mul(R(0.5), add(pdu, pdv)));
    const H3 du = mul((R)c.defocus_radius, u), dv = mul((R)c.defocus_radius, v);
    const H3 src[6] = {from, psl, pdu, pdv, du, dv};
    R* dst[6] = {rp.from, rp.psl, rp.pdu, rp.pdv, rp.du, rp.dv};
    for (int k = 0; k < 6; ++k) { dst[k][0] = src[k].x; dst[k][1] = src[k].y; dst[k][2] = src[k].z; }
}

// events created by one call: destroyed on every exit path (an early CRB_CUDA return used to leak them)
struct EventBag {
    std::vector<cudaEvent_t> ev;
    ~EventBag() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
    cudaError_t make(cudaEvent_t* e, unsigned flags = cudaEventDefault) {
        const cudaError_t r = cudaEventCreateWithFlags(e, flags);
        if (r == cudaSuccess) ev.push_back(*e);
        return r;
    }
};

struct EventTimer {
    bool on;
    cudaStream_t st;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans[4];
    ~EventTimer() {
        for (auto& v : spans)
            for (auto& p : v) {
                cudaEventDestroy(p.first);
                cudaEventDestroy(p.second);
            }
    }
    void begin(int cls, cudaEvent_t& a) {
        if (!on) return;
        cudaEventCreate(&a);
        cudaEventRecord(a, st);
        (void)cls;
    }
    void end(int cls, cudaEvent_t a) {
        if (!on) return;
        cudaEvent_t b;
        cudaEventCreate(&b);
        cudaEventRecord(b, st);
        spans[cls].push_back({a, b});
    }
    double total(int cls) {
        double ms = 0;
        for (auto& p : spans[cls]) {
            float f = 0;
            cudaEventElapsedTime(&f, p.first, p.second);
            ms += f;
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
        spans[cls].clear();
        return ms;
    }
};

// what the trace-variant choice depends on: the shape of the committed scene, not its coordinates
static uint64_t scene_signature(const SceneDeviceData& s, int precision) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    mix(s.n_nodes); mix(s.n_prims[0]); mix(s.n_prims[1]); mix(s.n_prims[2]); mix((uint64_t)precision); mix((uint64_t)s.sky_kind);
    return h;
}

template <typename R>
int render_impl(const SceneDeviceData& s, Workspace& ws, const CrCamera& cam_in, const CrRenderOpts& opts, void* d_out_rgb,
                void* d_out_rgb8, int packed, cudaStream_t stream, CrStats* stats, std::string& err) {
    const uint32_t W = cam_in.image_width, H = cam_in.image_height;
    const uint32_t world = opts.row_world <= 1 ? 1 : opts.row_world;
    const uint32_t block = opts.row_block == 0 ? 8 : opts.row_block;
    const uint32_t rank = world == 1 ? 0 : opts.row_rank;
    uint32_t rows_local = 0;
    if (world == 1) {
        rows_local = H;
    } else {
        for (uint32_t j = 0; j < H; ++j)
            if ((j / block) % world == rank) ++rows_local;
    }
    const uint64_t npix = (uint64_t)W * rows_local;
    const uint64_t total = npix * cam_in.samples;
    // few, large wavefronts: per-iteration launch gaps and kernel tails dominate below ~4 M paths, and throughput
    // still rises slowly to 64 M (book1 1080p x 100 spp: 1520 / 1525 / 1542 Msamples/s at 16 / 32 / 64 M, profiles/).
    // 64 M f64 paths = 16 GB of records + 6 GB of filter records + 2.5 GB of queues: 14 % of a B200's 180 GB.
    uint32_t pool = opts.pool_paths ? opts.pool_paths : (64u << 20);
    if (pool < 1024) pool = 1024;
    if ((uint64_t)pool > total && total > 0) pool = (uint32_t)((total + 31) & ~31ull);
    pool = (pool + 31u) & ~31u;  // whole warps; also keeps the hit arrays behind the queues 16 B aligned
    // u64 fixed point: 2^-44 resolution unless max_radiance * spp would overflow 62 bits
    uint32_t scale_bits = 44u;
    while (scale_bits > 8u && s.max_radiance * (double)cam_in.samples * (double)(1ull << scale_bits) >= 4.0e18) --scale_bits;
    const double fb_scale = (double)(1ull << scale_bits);

    // workspace carve-up
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t o_ctl = carve(sizeof(Control));
    const size_t o_cam = carve(sizeof(DevCamera));
    const size_t o_paths0 = carve((size_t)pool * sizeof(PathRec<R>));
    const size_t o_paths1 = carve((size_t)pool * sizeof(PathRec<R>));
    const size_t o_queues = carve((size_t)pool * Q_COUNT * sizeof(uint2) + (size_t)pool * HIT_QUEUES * sizeof(HitEntry<R>));  // queues, then their hits
    const size_t o_filt0 = carve((size_t)pool * sizeof(FilterRec));
    const size_t o_filt1 = carve((size_t)pool * sizeof(FilterRec));
    const size_t o_retry = carve((size_t)pool * sizeof(uint32_t));
    const size_t o_fb = carve((size_t)npix * 3 * sizeof(unsigned long long));
    int rc = ws.ensure(off, err);
    if (rc != CR_OK) return rc;
    char* base = static_cast<char*>(ws.ptr);
    Control* ctl = reinterpret_cast<Control*>(base + o_ctl);
    DevCamera* d_cam = reinterpret_cast<DevCamera*>(base + o_cam);
    PathRec<R>* paths[2] = {reinterpret_cast<PathRec<R>*>(base + o_paths0), reinterpret_cast<PathRec<R>*>(base + o_paths1)};
    uint2* queues = reinterpret_cast<uint2*>(base + o_queues);
    const HitEntry<R>* qhits = reinterpret_cast<const HitEntry<R>*>(queues + (size_t)Q_COUNT * pool);  // written by the trace kernels' commit
    FilterRec* filt[2] = {reinterpret_cast<FilterRec*>(base + o_filt0), reinterpret_cast<FilterRec*>(base + o_filt1)};
    unsigned long long* fb = reinterpret_cast<unsigned long long*>(base + o_fb);
    uint32_t* retry_list = reinterpret_cast<uint32_t*>(base + o_retry);

    if (!ws.pinned) {
        CRB_CUDA(cudaHostAlloc(&ws.pinned, 4096, cudaHostAllocDefault));
    }
    DevCamera hcam;
    memset(&hcam, 0, sizeof(hcam));
    hcam.c = cam_in;
    hcam.row_block = block;
    hcam.row_rank = rank;
    hcam.row_world = world;
    hcam.rows_local = rows_local;
    hcam.seed = opts.seed;
    hcam.fb_scale_bits = scale_bits;
    hcam.is_static = (cam_in.n_from_keys == 0 && cam_in.n_at_keys == 0) ? 1u : 0u;
    Control hctl;
    memset(&hctl, 0, sizeof(hctl));
    // ray_color(depth == 0) returns black BEFORE world.hit (ray_casting.rs:113-116): a max_depth 0 camera casts no ray
    // at all, the image is black and CrStats.rays is 0
    const uint64_t issue = cam_in.max_depth == 0 ? 0 : total;
    hctl.total_samples = issue;
    hctl.pool = pool;

    EventBag events;
    cudaEvent_t ev_begin, ev_end;
    CRB_CUDA(events.make(&ev_begin));
    CRB_CUDA(events.make(&ev_end));
    CRB_CUDA(cudaEventRecord(ev_begin, stream));
    // staging copies come from pinned memory so they are truly asynchronous
    char* pin = static_cast<char*>(ws.pinned);
    memcpy(pin, &hctl, sizeof(hctl));
    memcpy(pin + 512, &hcam, sizeof(hcam));
    static_assert(sizeof(Control) <= 512 && sizeof(DevCamera) <= 3072, "pinned staging layout");
    CRB_CUDA(cudaMemcpyAsync(ctl, pin, sizeof(hctl), cudaMemcpyHostToDevice, stream));
    CRB_CUDA(cudaMemcpyAsync(d_cam, pin + 512, sizeof(hcam), cudaMemcpyHostToDevice, stream));
    CRB_CUDA(cudaMemsetAsync(fb, 0, (size_t)npix * 3 * sizeof(unsigned long long), stream));

    const DevScene<R> sc = make_dev_scene<R>(s);
    // Two register budgets of the same trace kernel: 64 registers / 32 warps per SM and 48 registers / 40 warps per
    // SM.  The kernel is latency bound, so the extra warps win where the cheap node steps dominate (book1, meshes:
    // +4 %); where the f64 exact / leaf steps dominate (thin padded boxes of the Cornell quads) the spills of the
    // small budget lose 20 %.  Unless CRB_MINB pins one, the first large wavefront of a scene is traced with both
    // (same rays, same result) and the faster one is kept; the choice is cached per device by scene signature.
    typedef void (*TraceFn)(DevScene<R>, const PathRec<R>*, Control*, int, uint2*, const FilterRec*, uint32_t, const uint32_t*);
    const bool animated = s.anim_keys != nullptr;  // object keyframes: the builds that evaluate timelines at the ray time
    // Static scenes with a search tree run the order-free engine (fast_trace.cuh); the rays it hands back are traced by
    // the reference-order kernel in a second, small launch.  CR_RENDER_REFERENCE_ORDER keeps reference order throughout.
    const bool fast_ok = use_fast_engine(s, (opts.flags & CR_RENDER_REFERENCE_ORDER) != 0u);
    // f64: 72 registers / 7 CTAs = 28 warps per SM.  The box-test constants of a lane live in the lane table between INNER
    // slices (fast_trace.cuh), so this budget keeps the inner loop free of local-memory traffic; measured against the
    // 96-register / 5-CTA build: book1 -3 %, teapot -5 %, 10 M triangles -15 % trace time.  8 CTAs (64 registers, 200 KB of
    // lane tables, 56 KB of L1 left) lose 15 %.  f32: no build spills, the 8-CTA one (63 registers) has the most warps
    int fmb = sizeof(R) == 8 ? 7 : 8;
    if (const char* e = getenv("CRB_FAST_MINB")) fmb = atoi(e);
    auto fast_fn = fmb <= 3 ? k_trace_fast<R, 3> : fmb <= 4 ? k_trace_fast<R, 4> : fmb <= 5 ? k_trace_fast<R, 5> : fmb <= 6 ? k_trace_fast<R, 6> : fmb <= 7 ? k_trace_fast<R, 7> : k_trace_fast<R, 8>;
    const int g_fast = persistent_grid(fast_fn, TRACE_BLOCK, s.num_sms);
    TraceFn trace_variants[2] = {animated ? k_trace<R, CRB_REFILL, 8, true> : k_trace<R, CRB_REFILL, 8, false>,
                                 animated ? k_trace<R, CRB_REFILL, 8, true> : k_trace<R, CRB_REFILL, 10, false>};
    const int trace_grids[2] = {persistent_grid(trace_variants[0], TRACE_BLOCK, s.num_sms),
                                persistent_grid(trace_variants[1], TRACE_BLOCK, s.num_sms)};
    // Four ways to trace a wavefront, all with the reference's results: the reference-order kernel at two register budgets
    // (0: 64 registers, 1: 48 registers / more warps), the order-free engine (2) and its shared-memory build for small
    // scenes (3).  On every BASELINE config the order-free engine wins on mixed wavefronts (book1 1.3x, Cornell 1.3x, teapot
    // 1.5x, 10 M triangles 2.7x); only a wavefront of pure camera rays favours reference order, which made a one-wavefront
    // timing pick the wrong engine for short renders.  So: scenes with a search tree use the order-free engine (shared-memory
    // build when it fits) and fall back to reference order only if it hands back more than 5 % of its rays; scenes without
    // one (object keyframes, nested elements, CR_RENDER_REFERENCE_ORDER) time the two reference-order builds on their first
    // mixed wavefront (same rays, same result, k_trace_rewind between) and cache the choice per device by scene signature.
    // CRB_TRAVERSAL / CRB_MINB pin a candidate.
    const uint64_t signature = scene_signature(s, (sizeof(R) == 8 ? 0 : 1) | (fast_ok ? 2 : 0));
    int variant = ws.lookup_variant(signature);
    if (const char* e = getenv("CRB_MINB")) variant = atoi(e) <= 8 ? 0 : 1;
    if (const char* e = getenv("CRB_TRAVERSAL")) {
        if (e[0] == 'f' && fast_ok) variant = 2;
    }
    if (variant == 2 && !fast_ok) variant = -1;
    // (3: the order-free engine with the search tree in shared memory, for scenes small enough)
    // node records at a stride of 80 B when that still fits (conflict-free LDS.128 of a warp's divergent node fetches:
    // SmemTree::node_stride), else packed at 64 B.  CRB_NODE_STRIDE pins one (A/B).
    uint32_t node_stride = 80u;
    if (const char* e = getenv("CRB_NODE_STRIDE")) node_stride = atoi(e) == 64 ? 64u : 80u;
    if (fast_smem_layout<R>(s.n_fast_nodes, s.n_fast_prims, s.n_prims[0], s.n_prims[1], s.n_prims[2], node_stride).total > FAST_SMEM_LIMIT) node_stride = 64u;
    const size_t smem_need = fast_smem_layout<R>(s.n_fast_nodes, s.n_fast_prims, s.n_prims[0], s.n_prims[1], s.n_prims[2], node_stride).total;
    bool smem_ok = fast_ok && s.n_fast_nodes != 0u && smem_need <= FAST_SMEM_LIMIT;
    if (smem_ok && cudaFuncSetAttribute(k_trace_fast_smem<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need) != cudaSuccess) {
        cudaGetLastError();
        smem_ok = false;
    }
    if (const char* e = getenv("CRB_TRAVERSAL")) {
        if (e[0] == 's' && smem_ok) variant = 3;
    }
    if (variant == 3 && !smem_ok) variant = -1;
    int cands[2], n_cands = 0;
    cands[n_cands++] = 0;
    if (!animated) cands[n_cands++] = 1;
    if (fast_ok && variant < 0) variant = smem_ok ? 3 : 2;
    const uint64_t tune_at = total >= 2ull * pool ? 1 : 0;  // the second wavefront mixes bounce rays with camera rays
    bool tuning = variant < 0 && total >= (1ull << 20) && n_cands > 1;
    if (variant < 0) variant = fast_ok ? (smem_ok ? 3 : 2) : (animated ? 0 : 1);
    auto launch_trace = [&](int v, int cur) {
        if (v == 3) {
            k_trace_fast_smem<R><<<s.num_sms, FAST_BIG_BLOCK, smem_need, stream>>>(sc, paths[cur], ctl, cur, queues, filt[cur], pool, retry_list,
                                                                                  s.n_fast_nodes, s.n_fast_prims, s.n_prims[0], s.n_prims[1], s.n_prims[2], node_stride);
            trace_variants[0]<<<s.num_sms * 2, TRACE_BLOCK, 0, stream>>>(sc, paths[cur], ctl, cur, queues, filt[cur], pool, retry_list);
            return 2;
        }
        if (v == 2) {
            fast_fn<<<g_fast, TRACE_BLOCK, 0, stream>>>(sc, paths[cur], ctl, cur, queues, filt[cur], pool, retry_list);
            // the rays it could not decide, in reference order (usually none: the launch returns at once)
            trace_variants[0]<<<s.num_sms * 2, TRACE_BLOCK, 0, stream>>>(sc, paths[cur], ctl, cur, queues, filt[cur], pool, retry_list);
            return 2;
        }
        trace_variants[v]<<<trace_grids[v], TRACE_BLOCK, 0, stream>>>(sc, paths[cur], ctl, cur, queues, filt[cur], pool, nullptr);
        return 1;
    };
    int smb = 6;
    if (const char* e = getenv("CRB_SHADE_MINB")) smb = atoi(e);
    typedef void (*GenFn)(const Control*, const DevCamera*, const RaygenParams<R>, PathRec<R>*, FilterRec*);
    typedef void (*ScatFn)(DevScene<R>, const PathRec<R>*, PathRec<R>*, Control*, int, const uint2*, const HitEntry<R>*, FilterRec*, uint64_t, uint32_t);
    GenFn gen_fn = smb <= 4 ? k_raygen<R, 4> : smb <= 6 ? k_raygen<R, 6> : k_raygen<R, 8>;
    ScatFn lam_fn = animated ? k_shade_scatter<R, CR_MAT_LAMBERTIAN, 6, true>
                              : (smb <= 4 ? k_shade_scatter<R, CR_MAT_LAMBERTIAN, 4, false> : smb <= 6 ? k_shade_scatter<R, CR_MAT_LAMBERTIAN, 6, false> : k_shade_scatter<R, CR_MAT_LAMBERTIAN, 8, false>);
    ScatFn met_fn = animated ? k_shade_scatter<R, CR_MAT_METAL, 6, true>
                              : (smb <= 4 ? k_shade_scatter<R, CR_MAT_METAL, 4, false> : smb <= 6 ? k_shade_scatter<R, CR_MAT_METAL, 6, false> : k_shade_scatter<R, CR_MAT_METAL, 8, false>);
    ScatFn die_fn = animated ? k_shade_scatter<R, CR_MAT_DIELECTRIC, 6, true>
                              : (smb <= 4 ? k_shade_scatter<R, CR_MAT_DIELECTRIC, 4, false> : smb <= 6 ? k_shade_scatter<R, CR_MAT_DIELECTRIC, 6, false> : k_shade_scatter<R, CR_MAT_DIELECTRIC, 8, false>);
    RaygenParams<R> rp;
    memset(&rp, 0, sizeof(rp));
    rp.W = W; rp.rows_local = rows_local; rp.row_block = block; rp.row_rank = rank; rp.row_world = world;
    rp.is_static = hcam.is_static; rp.lens = !(cam_in.defocus_angle <= 0.0) ? 1u : 0u; rp.small = total < 0xFFFFFFFFull ? 1u : 0u;
    rp.seed = opts.seed;
    rp.bsmall = s.bsmall;
    rp.bmax = s.bmax;
    host_camera_constants<R>(cam_in, rp);
    const int g_gen = persistent_grid(gen_fn, SHADE_BLOCK, s.num_sms);
    auto tail_fn = animated ? k_tail<R, true, false> : k_tail<R, false, false>;
    auto tail_fast_fn = fast_ok ? k_tail<R, false, true> : tail_fn;  // the order-free engine's trace step inside the tail
    const int g_tail = std::min(persistent_grid(tail_fn, TRACE_BLOCK, s.num_sms), persistent_grid(tail_fast_fn, TRACE_BLOCK, s.num_sms));
    bool samples_out = total <= (uint64_t)pool;  // every camera sample issued by the prologue?
    uint32_t tail_n = 65536;
    if (const char* e = getenv("CRB_TAIL")) tail_n = (uint32_t)atoi(e);
    const int g_miss = persistent_grid(k_shade_miss<R>, SHADE_BLOCK, s.num_sms);
    const int g_emit = persistent_grid(k_shade_emissive<R>, SHADE_BLOCK, s.num_sms);
    const int g_lam = persistent_grid(lam_fn, SHADE_BLOCK, s.num_sms);
    const int g_met = persistent_grid(met_fn, SHADE_BLOCK, s.num_sms);
    const int g_die = persistent_grid(die_fn, SHADE_BLOCK, s.num_sms);

    EventTimer tm;
    tm.on = opts.time_kernels != 0;
    tm.st = stream;
    uint64_t launches = 0;
    // n_in of iteration k lands in pinned slot k % RING; the host looks LAG iterations behind so the
    // GPU always has work queued while the host decides whether the render is finished
    constexpr int RING = 8, LAG = 3;
    volatile uint32_t* h_n_in = reinterpret_cast<volatile uint32_t*>(pin + 3584);
    volatile uint32_t* h_retried = reinterpret_cast<volatile uint32_t*>(pin + 3584 + 64);  // retry_total as of the same iteration
    uint64_t traced_known = 0;
    uint32_t* d_n_in_ring = reinterpret_cast<uint32_t*>(pin + 3584);  // plan writes through zero-copy? no: device copy below
    (void)d_n_in_ring;
    cudaEvent_t ring_ev[RING];
    for (int i = 0; i < RING; ++i) CRB_CUDA(events.make(&ring_ev[i], cudaEventDisableTiming));

    cudaEvent_t a;
    // prologue: camera basis (static cameras), plan + raygen fill side 0
    tm.begin(2, a);
    k_plan<<<1, 32, 0, stream>>>(ctl, 0, nullptr, 0u);
    gen_fn<<<g_gen, SHADE_BLOCK, 0, stream>>>(ctl, d_cam, rp, paths[0], filt[0]);
    tm.end(2, a);
    launches += 2;

    uint64_t it = 0;
    bool done = (issue == 0);
    bool small_wave = false;
    const bool log_waves = getenv("CRB_LOG_WAVES") != nullptr;  // debug: wavefront sizes (pairs an ncu capture with its ray count)
    while (!done) {
        const int cur = (int)(it & 1), nxt = cur ^ 1;
        if (tuning && it == tune_at) {
            cudaEvent_t te[4];
            for (auto& e : te) CRB_CUDA(events.make(&e));
            for (int c = 0; c < n_cands; ++c) {
                CRB_CUDA(cudaEventRecord(te[2 * c], stream));
                launches += launch_trace(cands[c], cur);
                CRB_CUDA(cudaEventRecord(te[2 * c + 1], stream));
                if (c + 1 < n_cands) k_trace_rewind<<<1, 32, 0, stream>>>(ctl);
            }
            CRB_CUDA(cudaEventSynchronize(te[2 * n_cands - 1]));
            float best_ms = 0.f;
            for (int c = 0; c < n_cands; ++c) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, te[2 * c], te[2 * c + 1]);
                if (c == 0 || ms < best_ms) {
                    best_ms = ms;
                    variant = cands[c];
                }
            }
            ws.store_variant(signature, variant);
            tuning = false;
            launches += n_cands - 2;  // the rewinds; one trace launch is part of the five counted below
        } else {
            tm.begin(0, a);
            // (small wavefronts: the 43 us floor of the shared-memory build — tree copy, 896-thread CTAs — is not worth it)
            launches += launch_trace(variant == 3 && small_wave ? 2 : variant, cur) - 1;
            tm.end(0, a);
        }
        tm.begin(1, a);
        k_shade_miss<R><<<g_miss, SHADE_BLOCK, 0, stream>>>(sc, paths[cur], ctl, queues + (size_t)Q_MISS * pool, fb, fb_scale);
        lam_fn<<<g_lam, SHADE_BLOCK, 0, stream>>>(sc, paths[cur], paths[nxt], ctl, nxt, queues + (size_t)Q_LAMBERTIAN * pool, qhits, filt[nxt],
                                                  opts.seed, cam_in.max_depth);
        met_fn<<<g_met, SHADE_BLOCK, 0, stream>>>(sc, paths[cur], paths[nxt], ctl, nxt, queues + (size_t)Q_METAL * pool, qhits + (size_t)pool, filt[nxt],
                                                  opts.seed, cam_in.max_depth);
        die_fn<<<g_die, SHADE_BLOCK, 0, stream>>>(sc, paths[cur], paths[nxt], ctl, nxt, queues + (size_t)Q_DIELECTRIC * pool, qhits + 2 * (size_t)pool,
                                                  filt[nxt], opts.seed, cam_in.max_depth);
        launches += 5;
        if (!s.clamp_colors) {
            k_shade_emissive<R><<<g_emit, SHADE_BLOCK, 0, stream>>>(sc, paths[cur], ctl, queues + (size_t)Q_EMISSIVE * pool, fb,
                                                                    fb_scale);
            ++launches;
        }
        tm.end(1, a);
        tm.begin(2, a);
        k_plan<<<1, 32, 0, stream>>>(ctl, nxt, nullptr, tail_n);
        gen_fn<<<g_gen, SHADE_BLOCK, 0, stream>>>(ctl, d_cam, rp, paths[nxt], filt[nxt]);
        tm.end(2, a);
        launches += 2;
        if (samples_out && tail_n != 0u) {
            // When the pool is no longer full every camera sample has been issued; from then on each iteration offers
            // the wavefront to k_tail, which returns at once until k_plan finds at most tail_n paths left: then ONE launch
            // follows each of them to its end instead of ~25-45 more 8-launch iterations, and the iterations still in
            // flight find an empty wavefront.
            tm.begin(0, a);
            (variant >= 2 ? tail_fast_fn : tail_fn)<<<g_tail, TRACE_BLOCK, 0, stream>>>(sc, paths[nxt], ctl, nxt, opts.seed, cam_in.max_depth, fb, fb_scale);
            k_tail_finish<<<1, 32, 0, stream>>>(ctl, nxt);
            tm.end(0, a);
            launches += 2;
        }

        const int slot = (int)(it % RING);
        CRB_CUDA(cudaMemcpyAsync((void*)(h_n_in + slot), &ctl->n_in[nxt], sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        if (variant >= 2) CRB_CUDA(cudaMemcpyAsync((void*)(h_retried + slot), &ctl->retry_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CRB_CUDA(cudaEventRecord(ring_ev[slot], stream));
        ++it;
        if (it >= (uint64_t)LAG) {
            const int old = (int)((it - LAG) % RING);
            CRB_CUDA(cudaEventSynchronize(ring_ev[old]));
            const uint32_t left = h_n_in[old];
            if (log_waves) fprintf(stderr, "crucible_b200 wave %llu: %u paths enter iteration %llu\n", (unsigned long long)(it - LAG),
                                   left, (unsigned long long)(it - LAG + 1));
            if (left < pool) samples_out = true;
            small_wave = left < (1u << 18);
            traced_known += left;
            // safety valve: a scene whose rays the order-free engine keeps handing back (irregular rays: exact zero direction
            // components, e.g. axis-parallel beams between axis-aligned mirrors) is better off in reference order throughout
            if (variant >= 2 && traced_known >= (1u << 16) && (uint64_t)h_retried[old] * 20u > traced_known) variant = animated ? 0 : 1;
            if (left == 0) done = true;  // nothing left to trace after iteration it-LAG (k_tail took over, or every path ended)
        }
        if (it > 100000000ull) {
            err = "render: iteration limit";
            return CR_ERR_LIMIT;
        }
    }
    CRB_CUDA(cudaGetLastError());
    // resolve
    cudaEvent_t r0;
    tm.begin(3, r0);
    if (npix > 0) {
        int g_res = (int)((npix + 255) / 256);
        if (g_res > s.num_sms * 8) g_res = s.num_sms * 8;
        k_resolve<<<g_res, 256, 0, stream>>>(fb, (uint32_t)npix, W, block, rank, world, packed, 1.0 / fb_scale,
                                             (double)cam_in.samples, s.clamp_colors ? 0 : 1, static_cast<double*>(d_out_rgb),
                                             static_cast<uint8_t*>(d_out_rgb8));
        ++launches;
    }
    tm.end(3, r0);
    CRB_CUDA(cudaMemcpyAsync(pin, ctl, sizeof(Control), cudaMemcpyDeviceToHost, stream));
    CRB_CUDA(cudaEventRecord(ev_end, stream));
    CRB_CUDA(cudaEventSynchronize(ev_end));
    CRB_CUDA(cudaGetLastError());
    if (stats) {
        Control fin;
        memcpy(&fin, pin, sizeof(fin));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        stats->samples = total;
        stats->rays = fin.rays_traced;
        stats->retried_rays = (uint64_t)fin.retry_total + fin.retry_count;
        stats->trace_engine = (uint32_t)variant;
        stats->iterations = it;
        stats->launches = launches;
        stats->ms_total = ms;
        stats->ms_trace = tm.total(0);
        stats->ms_shade = tm.total(1);
        stats->ms_raygen = tm.total(2);
        stats->ms_resolve = tm.total(3);
    } else {
        for (int c = 0; c < 4; ++c) tm.total(c);
    }
    return CR_OK;
}

}  // namespace crb
