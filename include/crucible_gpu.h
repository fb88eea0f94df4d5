/*
 * crucible_gpu.h — C ABI of the B200-native path-tracing backend for Crucible.
 *
 * This is the drop-in boundary: the entry points below are exactly what a
 * `mod gpu` inside the Crucible crate would bind over `extern "C"` to replace the
 * CPU sample loop of `Camera::render` (reference src/camera/mod.rs:270-317, the
 * part between the PPM header and the write loop, :284-303).  Every function cites
 * the reference item it replaces.  All citations are paths under the reference
 * repository (kylittle/Crucible).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types in any signature;
 *   - every function returns 0 on success or a negative CrStatus; it never aborts
 *     (the reference panics instead, e.g. src/camera/mod.rs:294,302,308);
 *     `cr_last_error()` returns a thread-local message for the last failure;
 *   - inputs are borrowed for the duration of the call (the library copies what it
 *     needs to the device), outputs are caller-allocated;
 *   - one host thread drives one CrScene; a CrScene lives on one CUDA device
 *     (one process per GPU; multi-GPU jobs shard image rows or frames, see
 *     CrRenderOpts.row_*);
 *   - there is NO CPU fallback: with no CUDA device every compute entry point
 *     fails with CR_ERR_NO_DEVICE.
 *
 * Numeric contract
 *   - CR_PRECISION_F64 reproduces the reference's f64 arithmetic operation by
 *     operation (no FMA contraction, reference traversal order): closest-hit
 *     `prim_index` and `front_face` are bit-exact, t / normal / uv differ only
 *     through libm (acos/atan2/asin) by a few ulp.
 *   - CR_PRECISION_F32 is the fast path (the same reference-order traversal on f32
 *     boxes and primitives, FMA allowed); it is statistically equivalent, not bit-exact.
 */
#ifndef CRUCIBLE_GPU_H
#define CRUCIBLE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CrScene CrScene; /* opaque; owns host staging + device copies */

typedef enum CrStatus {
    CR_OK = 0,
    CR_ERR_INVALID = -1,    /* bad argument (null pointer, index out of range, NaN where forbidden) */
    CR_ERR_NO_DEVICE = -2,  /* no CUDA device / wrong architecture: there is no CPU fallback */
    CR_ERR_CUDA = -3,       /* a CUDA runtime call failed; see cr_last_error() */
    CR_ERR_STATE = -4,      /* call order violated (e.g. render before cr_scene_commit) */
    CR_ERR_LIMIT = -5       /* a built-in limit was exceeded (BVH depth, texture nesting) */
} CrStatus;

/* ---- materials: reference src/materials/mod.rs:16-20 (+ Emissive extension) ---- */
typedef enum CrMaterialKind {
    CR_MAT_LAMBERTIAN = 0, /* src/materials/lambertian.rs:40-61 */
    CR_MAT_METAL = 1,      /* src/materials/metal.rs:29-42 */
    CR_MAT_DIELECTRIC = 2, /* src/materials/dielectric.rs:21-55 */
    CR_MAT_EMISSIVE = 3    /* EXTENSION (absent in reference): emits `emit`, never scatters */
} CrMaterialKind;

typedef struct CrMaterial {
    int32_t kind;        /* CrMaterialKind */
    int32_t tex;         /* Lambertian: texture index (Lambertian.tex, lambertian.rs:17) */
    double scatter_prob; /* Lambertian.scatter_prob (lambertian.rs:18); must be > 0 */
    double albedo[3];    /* Metal.albedo (metal.rs:12) */
    double fuzz;         /* Metal.fuzz (metal.rs:13), in [0,1] */
    double ior;          /* Dielectric.refraction_index (dielectric.rs:10) */
    double emit[3];      /* Emissive radiance (extension; unclamped) */
} CrMaterial;

/* ---- textures: reference src/textures/mod.rs:13-17 ---- */
typedef enum CrTextureKind {
    CR_TEX_SOLID = 0,   /* src/textures/solid_color.rs:25-27 */
    CR_TEX_CHECKER = 1, /* src/textures/checker_texture.rs:39-51 (even/odd are texture indices) */
    CR_TEX_IMAGE = 2    /* src/textures/image_texture.rs:23-32 (nearest texel of an RGB8 image) */
} CrTextureKind;

typedef struct CrTexture {
    int32_t kind;     /* CrTextureKind */
    int32_t even;     /* checker: texture index used when floor-sum is even */
    int32_t odd;      /* checker: texture index used when floor-sum is odd */
    int32_t image;    /* image: index returned by cr_scene_add_image */
    double color[3];  /* solid: albedo, each in [0,1] */
    double inv_scale; /* checker: 1.0/scale, computed by the caller as checker_texture.rs:24 does */
} CrTexture;

/* ---- sky: reference src/scene/mod.rs:19-26, src/camera/ray_casting.rs:133-151 ---- */
typedef enum CrSkyKind {
    CR_SKY_DEFAULT = 0,   /* white -> (0.5,0.7,1.0) lerp, ray_casting.rs:145-150 */
    CR_SKY_SPHERICAL = 1, /* equirect image, nearest texel, ray_casting.rs:134-144 + scene/mod.rs:37-45 */
    CR_SKY_BLACK = 2      /* EXTENSION for closed emissive scenes (Cornell box) */
} CrSkyKind;

/* ---- primitives ---- */
typedef enum CrPrimKind {
    CR_PRIM_SPHERE = 0,   /* src/objects/sphere.rs */
    CR_PRIM_TRIANGLE = 1, /* src/objects/triangle.rs */
    CR_PRIM_QUAD = 2      /* EXTENSION: parallelogram Q,u,v with a padded bbox */
} CrPrimKind;

/* ---- camera: reference src/camera/mod.rs:66-100 flattened ---- */
typedef enum CrInterp { CR_NERP = 0, CR_LERP = 1 } CrInterp; /* src/timeline/mod.rs:100-104 */

/* One translate keyframe of a camera timeline, already in the relative form the
 * reference stores (src/timeline/transform_builder.rs:348-469): at time t it
 * contributes  delta * s  on `axis`, with s = clamp((t-t0)/(t1-t0),0,1) for LERP and
 * 1 for NERP, and only when t >= t0 (Interval::is_less || contains,
 * src/timeline/mod.rs:239-243).  Keyframes are listed in the timeline's sorted order. */
typedef struct CrKeyframe {
    double t0, t1, delta;
    int32_t axis;   /* 0,1,2 */
    int32_t interp; /* CrInterp */
} CrKeyframe;

#define CR_MAX_CAM_KEYS 32

typedef struct CrCamera {
    uint32_t image_width, image_height; /* Viewport, camera/mod.rs:26-47 */
    double viewport_width, viewport_height; /* fix_viewport, rendering_compute.rs:5-11 (host computes tan) */
    double focus_dist;                      /* camera/mod.rs:81 */
    double defocus_angle;                   /* radians; <= 0 disables the thin lens (ray_casting.rs:96) */
    double defocus_radius;                  /* focus_dist * tan(defocus_angle/2), rendering_compute.rs:72-74 */
    double vup[3];
    double look_from[3], look_at[3];        /* InitTranslate of the two timelines */
    double frame_rate;                      /* camera/mod.rs:95 */
    uint32_t frame;                         /* camera/mod.rs:96 */
    uint32_t samples;                       /* camera/mod.rs:84 */
    uint32_t max_depth;                     /* camera/mod.rs:86 */
    uint32_t n_from_keys, n_at_keys;
    double shutter_angle;                   /* degrees, camera/mod.rs:99 */
    CrKeyframe from_keys[CR_MAX_CAM_KEYS];  /* cam_translate_*("from"), scene_animator.rs:460-552 */
    CrKeyframe at_keys[CR_MAX_CAM_KEYS];
} CrCamera;

typedef enum CrPrecision { CR_PRECISION_F64 = 0, CR_PRECISION_F32 = 1 } CrPrecision;

typedef struct CrRenderOpts {
    uint64_t seed;        /* Philox key; same seed => same image, for any GPU count */
    int32_t precision;    /* CrPrecision */
    uint32_t pool_paths;  /* wavefront pool size; 0 = default */
    /* Row sharding for multi-GPU: this call renders only rows j with
     * (j / row_block) % row_world == row_rank.  row_world = 0 or 1 renders everything.
     * The RNG is keyed by the GLOBAL pixel index, so the union over ranks is
     * bit-identical to a single-GPU render. */
    uint32_t row_block, row_rank, row_world;
    uint32_t time_kernels; /* 1 = bracket every kernel class with CUDA events (CrStats.ms_*) */
    uint32_t flags;        /* CR_RENDER_* bits */
} CrRenderOpts;
/* cr_render_device only: d_out_rgb / d_out_rgb8 are FULL [H][W][3] images and this rank's rows are written at their
 * global position (default: the rank's rows packed, [rows_local][W][3]).  The image may live on ANOTHER device
 * (peer-mapped memory, or a buffer opened with cr_shared_buffer_open): the resolve kernel then stores its rows
 * straight into that device's memory over NVLink, which is the whole framebuffer gather of SURVEY 8e. */
#define CR_RENDER_GLOBAL_ROWS 1u
/* Closest hits in the reference's own traversal order throughout (BVHWrapper::hit, bvhwrapper.rs:97-126: DFS over the
 * reference tree).  Default for static scenes is the order-free engine: the same hits found by a near-first search over
 * a tree of its own, every candidate confirmed against its reference leaf-node box, undecidable rays re-traced in
 * reference order (DESIGN.md 5.1b states the one numerical assumption this rests on).  Scenes with object keyframes
 * always use reference order. */
#define CR_RENDER_REFERENCE_ORDER 2u

typedef struct CrStats {
    uint64_t samples;      /* camera samples generated (W*H*spp over the rows rendered) */
    uint64_t rays;         /* ray segments traced = world.hit calls (primary + bounces) */
    uint64_t iterations;   /* wavefront iterations */
    uint64_t launches;     /* CUDA kernels launched by this call */
    double ms_total;       /* device time of the whole call (CUDA events on the render stream) */
    double ms_trace;       /* time_kernels=1: sum over trace launches */
    double ms_shade;       /* time_kernels=1: shade + miss kernels */
    double ms_raygen;      /* time_kernels=1: plan + raygen */
    double ms_resolve;
    double ms_h2d, ms_d2h; /* host-buffer entry points only */
    uint64_t retried_rays; /* order-free engine: ray segments handed back to the reference-order kernel */
    uint32_t trace_engine; /* what traced the wavefronts: 0 / 1 = reference-order kernel at 64 / 48 registers, 2 = order-free engine,
                              3 = order-free engine with the search tree in shared memory (small scenes) */
    uint32_t reserved;
} CrStats;

/* Result of a closest-hit query: the fields of HitRecord (src/objects/mod.rs:21-29)
 * plus the ids defined in SURVEY 8b. prim_index = insertion index in the scene's flat
 * element list (hidden primitives keep their slot); -1 = miss. */
typedef struct CrHit {
    int32_t prim_index;
    int32_t obj_id;
    int32_t front_face;
    int32_t material;
    double t;
    double p[3];
    double n[3];
    double u, v;
} CrHit;

/* ---- lifetime ---- */
/* Number of usable sm_100 devices (0 when CUDA is unavailable). */
int cr_device_count(void);
CrScene* cr_scene_create(int device);
/* Releases what the library caches on `device` between calls: the wavefront arena shared by the scenes of the process
 * (path pool, queues, framebuffer: up to ~25 GB at the default pool of 64 M paths), its pinned staging block, and the
 * blocks held by the device's stream-ordered memory pool, and the host staging blocks cached between scenes
 * (cr_scene_reserve).  Safe to call between renders (it waits for the device);
 * the next render allocates again.  The reference has nothing to release: its worker pool dies with Camera::render
 * (camera/mod.rs:299-303). */
int cr_device_trim(int device);
void cr_scene_destroy(CrScene*);
const char* cr_last_error(void);
const char* cr_version(void);

/* ---- scene description (replaces crate-private access to Scene.elements,
 *      src/scene/mod.rs:75-82, 159-230).  Each call APPENDS to the flat element list in
 *      call order, exactly like Scene::add_element / load_asset, and returns the index of
 *      the first appended primitive (>= 0) or a negative CrStatus. ---- */
/* Optional size hint before a run of cr_scene_add_* calls: room for this many MORE primitives of each kind, so that the
 * staging arrays grow once (Scene::load_asset adds one mesh per call, scene/mod.rs:211-229: a 10 M-triangle world arrives
 * as ~1 600 calls).  The reference's Vec::push has no counterpart of its own; a flattener that walks a finished
 * Hittables tree knows the totals.  Staging blocks of 1 MB and more are cached by the library between scenes (a per-frame
 * rebuild, Scene::render_image scene/mod.rs:332-347, then touches mapped memory); cr_device_trim releases them. */
int cr_scene_reserve(CrScene*, size_t n_spheres, size_t n_triangles, size_t n_quads);
/* Sphere::new, sphere.rs:25-39: cxyz_r = [n][4] (centre, radius >= 0). */
int64_t cr_scene_add_spheres(CrScene*, const double* cxyz_r, const int32_t* material,
                             const int32_t* obj_id, size_t n);
/* Triangle::new, triangle.rs:23-46: abc = [n][9]. */
int64_t cr_scene_add_triangles(CrScene*, const double* abc, const int32_t* material,
                               const int32_t* obj_id, size_t n);
/* EXTENSION: quads Q,u,v = [n][9]. */
int64_t cr_scene_add_quads(CrScene*, const double* quv, const int32_t* material,
                           const int32_t* obj_id, size_t n);
/* Many add calls in one: batch b appends counts[b] primitives of kind kinds[b] (CR_PRIM_SPHERE: data[b] = [n][4],
 * CR_PRIM_TRIANGLE / CR_PRIM_QUAD: [n][9]) exactly as the matching cr_scene_add_* call would, in order; material / obj_id may be
 * NULL, and so may their entries.  A flattener that walks Scene.elements (scene/mod.rs:75-82) has one batch per mesh
 * (load_asset, scene/mod.rs:211-229); together they are validated and copied on all host threads.  Returns the index of the
 * first appended primitive; on error nothing is appended. */
int64_t cr_scene_add_batches(CrScene*, size_t n_batches, const int32_t* kinds, const double* const* data,
                             const int32_t* const* material, const int32_t* const* obj_id, const size_t* counts);
/* ---- nested elements: Scene::add_element also takes a whole Hittables::HitList or Hittables::BVHWrapper as ONE element
 *      (scene/mod.rs:160-166).  Primitives added between cr_scene_begin_group and the matching cr_scene_end_group are the
 *      members of that element, in call order; groups nest.  Members keep flat prim_index values in call order.
 *        CR_GROUP_HITLIST : HitList built with HitList::add (hitlist.rs:27-30): hit() scans the members in order with one
 *                           shrinking interval and no box test (hitlist.rs:52-65); a hidden member never hits
 *                           (sphere.rs:62, triangle.rs:87) but still widens the list's box.
 *        CR_GROUP_BVH     : BVHWrapper::new_wrapper over the members (bvhwrapper.rs:15-44): hidden primitives are dropped,
 *                           the rest get their own median-split tree, which the enclosing tree meets as one leaf.
 *      The enclosing BVH sorts and boxes a group like any other element (its bounding_box()).  Scenes with groups use the
 *      reference-order engine and the host BVH builder.  cr_scene_begin_group returns the group id (>= 0). ---- */
typedef enum CrGroupKind { CR_GROUP_HITLIST = 0, CR_GROUP_BVH = 1 } CrGroupKind;
int cr_scene_begin_group(CrScene*, int kind);
int cr_scene_end_group(CrScene*);
/* Scene::hide_element / show_element, scene/mod.rs:232-281 (per primitive). */
int cr_scene_set_hidden(CrScene*, size_t prim_index, int hide);
/* At most 2^24 materials (CR_ERR_LIMIT at commit beyond that). */
int cr_scene_set_materials(CrScene*, const CrMaterial*, size_t n);
int cr_scene_set_textures(CrScene*, const CrTexture*, size_t n);
/* RTWImage (asset_loader/img_loader.rs:9-13): rgb8 = [h][w][3]; returns the image index. */
int cr_scene_add_image(CrScene*, const uint8_t* rgb8, int w, int h);
/* Scene::load_default_skybox / load_spherical_skybox, scene/mod.rs:146-156. */
int cr_scene_set_sky(CrScene*, int kind, int image);
/* ---- object animation: the keyframes of a primitive's TransformTimeline ---- */
/* One keyframe of src/timeline/transform_builder.rs in evaluated form.  At ray time t a key is VALID when
 * t > t1 or t0 <= t <= t1 (Interval::is_less || contains, timeline/mod.rs:239-243, 251-253), with
 * s = clamp((t - t0) / (t1 - t0), 0, 1) (Transform::get_matrix_at_time, timeline/mod.rs:88-96):
 *   kind 0,1,2 (translate_x/y/z, transform_builder.rs:348-713): every valid key ADDS  a*s (LERP) or a (NERP) to
 *              that axis, in list order (the product of translate matrices, timeline/mod.rs:244-248);
 *   kind 3..6  (scale keys): only the LAST valid scale key of the list counts (timeline/mod.rs:251-257); its value
 *              is v = a + (b - a)*s (LERP) or b (NERP) and its matrix multiplies the translated point (x, y, z):
 *       3 = scale_sphere (transform_builder.rs:18-96)   diag(1,1,1,v): the radius becomes v (spheres only);
 *       4 = scale_x (:101-180)  diag(v,1,1,1): x' = v*x            (triangle vertices only, scene_animator.rs:38-41)
 *       5 = scale_y (:186-265)  the reference writes v into row 1, column 0 (:229-246): y' = v*x + y, reproduced;
 *       6 = scale_z (:271-346)  diag(1,1,v,1): z' = v*z.
 *     With no valid scale key the construction radius / the identity applies.  Because only the last valid key
 *     counts, scale_point / scale_all_uniform (:729-733, scene_animator.rs:187-219), which push an x, a y and a z key
 *     with the same interval, end up scaling z only: reference behaviour, reproduced. */
typedef struct CrAnimKey {
    double t0, t1, a, b;
    int32_t kind;   /* 0,1,2 = translate axis; 3 = sphere radius; 4,5,6 = scale x, y, z */
    int32_t interp; /* CrInterp */
} CrAnimKey;
/* Replaces the keyframes of one point of a primitive: point 0 of a sphere (centre + radius), points 0,1,2 = the
 * a, b, c vertex timelines of a triangle (triangle.rs:14-16).  Keys in the timeline's sorted order (the stable
 * sort by interval start the builder applies after every insertion).  As in the reference, the BVH is built from
 * the construction-time boxes (bvhwrapper.rs:47-50: `update_bb` results are never read), so a primitive that
 * leaves its first box can only be hit through that box.  Quads (extension) are not animated. */
int cr_scene_set_keyframes(CrScene*, size_t prim_index, int point, const CrAnimKey* keys, size_t n);

/* Where cr_scene_commit builds the BVH.  Both builders produce the SAME tree as
 * BVHWrapper::help_generate (bvhwrapper.rs:46-94): same nodes in the same preorder, boxes bit for bit.
 *   HOST   : the recursion as written (stable sort per span), host threads for the top levels;
 *   DEVICE : level-synchronous build on the scene's GPU (one stable radix sort per tree level);
 *   AUTO   : DEVICE for scenes with a device and >= 32768 visible primitives, HOST otherwise.
 * (box_compare's NaN -> Equal branch, bvhwrapper.rs:92, is unreachable: non-finite coordinates are rejected
 * when primitives are added.)
 * CR_BVH_DEVICE on a scene without a device fails with CR_ERR_NO_DEVICE at commit. */
typedef enum CrBvhBuilder { CR_BVH_AUTO = 0, CR_BVH_HOST = 1, CR_BVH_DEVICE = 2 } CrBvhBuilder;
int cr_scene_set_bvh_builder(CrScene*, int builder);

/* BVHWrapper::new_wrapper (bvhwrapper.rs:15-94) reproduced (see CrBvhBuilder), then flattened
 * and uploaded.  Must be called after the last edit and before trace/render. */
int cr_scene_commit(CrScene*);
/* Wall-clock breakdown of the last cr_scene_commit (milliseconds). */
typedef struct CrCommitInfo {
    int32_t builder;   /* CR_BVH_HOST or CR_BVH_DEVICE: the builder that ran */
    uint32_t levels;   /* tree depth */
    double ms_total;   /* whole commit */
    double ms_build;   /* BVH build (DEVICE: pack + h2d + device + d2h below) */
    double ms_pack, ms_h2d, ms_device, ms_d2h; /* DEVICE builder phases; 0 for HOST */
    double ms_upload;  /* flatten to device records + H2D of the scene */
    double ms_search_tree; /* build + upload of the order-free engine's search tree (0 when the scene has none) */
} CrCommitInfo;
int cr_scene_commit_info(const CrScene*, CrCommitInfo* out);
/* One node of the committed BVH in preorder (tests compare the two builders bit for bit). */
typedef struct CrBvhNode {
    double lo[3], hi[3];  /* Aabb of the node (bvh.rs:19-23) */
    uint32_t left, right; /* inner node: preorder indices of the children; leaf node: bit31 | kind << 29 | index
                             of the primitive within its kind (right = 0x7FFFFFFF for a span-1 node) */
    uint32_t axis;        /* longest axis of the box = sort axis (bvhwrapper.rs:52) */
    uint32_t skip;        /* preorder index of the first node after this subtree */
} CrBvhNode;
/* Copies min(cap, n_nodes) nodes; returns n_nodes. */
int64_t cr_scene_bvh_nodes(const CrScene*, CrBvhNode* out, size_t cap);
/* The flattened records the trace kernels read, copied back from the device (tests compare the host and the
 * device flattening byte for byte).  which: 0 = nodes f64 (64 B each: xmin,xmax,ymin,ymax,zmin,zmax, two words,
 * padding), 1 = nodes f32 (32 B: the boxes rounded outward, the same two words), 2 = triangles f64 (80 B:
 * a, e1 = b-a, e2 = c-a, padding), 3 = triangles f32 (48 B).  Copies min(cap_bytes, total); returns total bytes. */
int64_t cr_scene_device_records(const CrScene*, int which, void* out, size_t cap_bytes);
/* Introspection of the committed BVH (tests): node count, max depth, leaf order. */
int cr_scene_bvh_info(const CrScene*, uint64_t* n_nodes, uint32_t* max_depth, uint64_t* n_visible);
/* DFS leaf order of the committed BVH as prim_index values (cap entries at most). */
int64_t cr_scene_bvh_leaf_order(const CrScene*, int32_t* out, size_t cap);

/* ---- the hot path ---- */
/* Replaces Hittables::hit (src/objects/mod.rs:118-125) on fixed ray batches.
 * rays = [n][7] (origin, direction, time); interval (tmin,tmax) open as Interval::surrounds.
 * The ray time positions animated primitives (Sphere::hit / Triangle::hit evaluate their timelines at r.time()). */
int cr_trace_batch(CrScene*, const double* rays, size_t n, double tmin, double tmax,
                   int precision, CrHit* out);
/* `precision` of cr_trace_batch = CrPrecision, optionally OR-ed with this bit: reference traversal order throughout
 * (see CR_RENDER_REFERENCE_ORDER). */
#define CR_TRACE_REFERENCE_ORDER 0x100
/* Rays of the last cr_trace_batch on this scene that the order-free engine handed back to the reference-order kernel. */
int64_t cr_scene_last_retried(const CrScene*);

/* Replaces thread_setup .. join of Camera::render (camera/mod.rs:284-303): one averaged
 * linear colour per pixel, row-major j then i (camera/mod.rs:306-311).
 *   out_rgb  : [H][W][3] f64 linear mean (average_samples, ray_casting.rs:154-173), or NULL
 *   out_rgb8 : [H][W][3] bytes floor(255*sqrt(c)) (Display for Color, utils.rs:422-438), or NULL
 * With row sharding only the rows of this rank are written; the others are left untouched. */
int cr_render(CrScene*, const CrCamera*, const CrRenderOpts*, double* out_rgb, uint8_t* out_rgb8,
              CrStats* stats);
/* Same, but the outputs are DEVICE pointers on the scene's device and the work is enqueued
 * on `cuda_stream` (a cudaStream_t; NULL = the library's own stream).  The call returns
 * after the stream has drained.  Used by the multi-GPU driver, which gathers rows over NCCL. */
int cr_render_device(CrScene*, const CrCamera*, const CrRenderOpts*, void* d_out_rgb,
                     void* d_out_rgb8, void* cuda_stream, CrStats* stats);

/* ---- multi-GPU behind the boundary (the `cr_init` over a device list, and 8e) ----
 * Stills shard by interleaved row blocks, the scene is replicated, and the only exchange is the framebuffer: every
 * device's resolve kernel writes its rows directly into the image on the first device (peer stores over NVLink /
 * NVSwitch; block copies when two devices cannot map each other), followed by ONE device-to-host copy.  The RNG is
 * keyed by the global pixel index, so the image is bit-identical to a single-device cr_render. */
/* A committed copy of `src` on another device (same primitives, materials, images, sky, keyframes, hidden flags and
 * BVH builder; the copy builds its own tree, which is the same tree).  NULL on failure. */
CrScene* cr_scene_replicate(const CrScene* src, int device);
/* Camera::render's sample loop over n devices from ONE host thread's point of view (the call drives one worker
 * thread per device).  replicas[i] must be committed copies of one scene on n DIFFERENT devices; replicas[0]'s
 * device assembles the image.  opts->row_rank / row_world are ignored (set per replica); opts->row_block = rows per
 * interleaved block (0 = 8).  out_rgb / out_rgb8 as for cr_render (either may be NULL).  stats: n entries or NULL
 * (entry i = replica i; ms_d2h of entry 0 = the final copy). */
int cr_render_multi(CrScene* const* replicas, int n, const CrCamera*, const CrRenderOpts*, double* out_rgb,
                    uint8_t* out_rgb8, CrStats* stats);
/* One process per GPU (torchrun-style launch): the rank that assembles the image creates a device buffer other
 * processes can map, ships the 64-byte handle to them by any means (a torch.distributed broadcast in
 * crucible_b200.multigpu), and every rank renders with CR_RENDER_GLOBAL_ROWS into it through cr_render_device.
 * create: cudaMalloc + cudaIpcGetMemHandle on `device`; open: cudaIpcOpenMemHandle in the calling process for use
 * from `device`; close: owner != 0 frees the allocation, owner == 0 unmaps it. */
int cr_shared_buffer_create(int device, size_t bytes, void** dptr, unsigned char handle[64]);
int cr_shared_buffer_open(int device, const unsigned char handle[64], void** dptr);
int cr_shared_buffer_close(int device, void* dptr, int owner);

/* ---- the file-writing tail of the path ---- */
typedef enum CrImageFormat {
    CR_PPM_P3 = 0, /* the reference's output: text PPM, camera/mod.rs:286,306-311 */
    CR_PPM_P6 = 1, /* EXTENSION: binary PPM, same header fields, 3 bytes per pixel */
    CR_PNG = 2     /* EXTENSION: 8-bit RGB PNG (zlib deflate, filter 0), the same bytes as the PPM */
} CrImageFormat;

/* Writes [h][w][3] bytes exactly as Camera::render does (camera/mod.rs:275-311): the file is
 * created or truncated, "P3\n{w} {h}\n255\n", then one "{r} {g} {b}\n" line per pixel, row-major.
 * Formatting is table driven and split over host threads (the reference formats 2 M `Display`
 * calls through a BufWriter). */
int cr_write_ppm(const char* path, const uint8_t* rgb8, uint32_t w, uint32_t h, int format);

/* ---- scene export file ----
 * The reference can only hand a scene to its renderer inside one process (Scene -> BVHWrapper -> Camera::render,
 * scene/mod.rs:332-347).  cr_scene_save writes everything a CrScene was given — primitives in insertion order with their
 * materials, ids and hidden flags, material / texture / image tables, sky, object keyframes, BVH builder choice — to one
 * little-endian binary file ("CRSCENE1" + counted sections); cr_scene_load rebuilds and COMMITS the scene on `device`
 * (-1 = host only).  A scene saved, loaded and rendered gives the image of the original, bit for bit. */
int cr_scene_save(const CrScene*, const char* path);
CrScene* cr_scene_load(const char* path, int device);

/* Camera::render(&skybox, &world, fname) as a whole (camera/mod.rs:270-317): sample loop on the
 * GPU, bytes to the host, file written.  Honors row sharding only when row_world <= 1. */
int cr_render_to_file(CrScene*, const CrCamera*, const CrRenderOpts*, const char* path, int format,
                      CrStats* stats);

/* The frame loop of Scene::render_movie (scene/mod.rs:295-322) for a static world: renders the
 * frames first, first+stride, ... (< n_frames) with Camera.frame = cam->frame + that index
 * (Camera::next_frame, camera/mod.rs:160-162) into `<dir>/image{frame:0>digits}.ppm` (scene/mod.rs:307-311).
 * `stride`/`first` shard whole frames over ranks.  The loop is pipelined: while frame k+1
 * is traced, frame k is copied to pinned host memory on a second stream, formatted and written by
 * a writer thread.  stats = array of ceil((n_frames-first)/stride) entries, or NULL. */
int cr_render_frames(CrScene*, const CrCamera* cam, const CrRenderOpts*, uint32_t first, uint32_t stride,
                     uint32_t n_frames, const char* dir, uint32_t digits, int format, CrStats* stats);

/* ---- host helpers on the path ---- */
/* TransformTimeline::combine_and_compute for a camera point (timeline/mod.rs:233-263):
 * out = init + sum of keyframe contributions at time t. */
int cr_camera_point_at(const double init[3], const CrKeyframe* keys, size_t n, double t, double out[3]);
/* TransformTimeline::combine_and_compute for an OBJECT point (timeline/mod.rs:233-263): init = (x, y, z, w) with
 * w = the construction radius of a sphere or 1.0 for a triangle vertex (TransformTimeline::new_sphere / ::new,
 * timeline/mod.rs:129-223); keys as for cr_scene_set_keyframes; out = (x, y, z, w) at time t.  The host copy of the
 * routine Sphere::hit / Triangle::hit evaluate on the device (sphere.rs:67-70, triangle.rs:91-97). */
int cr_anim_point_at(const double init[4], const CrAnimKey* keys, size_t n, double t, double out[4]);
/* FP64 / FP32 FMA peak of the device measured with a register-resident micro-kernel
 * (roofline denominators that MEASURED_PEAKS.json does not hold). TFLOP/s. */
int cr_measure_fma_peak(int device, double* fp64_tflops, double* fp32_tflops);
/* Philox4x32-10 block (tests pin it against the Random123 known answers). */
void cr_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* CRUCIBLE_GPU_H */
