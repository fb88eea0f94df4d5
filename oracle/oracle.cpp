/*
 * oracle.cpp — CPU restatement of Crucible's path-tracing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under crucible_b200/ links, imports or calls it.
 *
 * What it is: a function-by-function C++17 restatement of the reference's f64 algorithm
 * (kylittle/Crucible, Rust; no rustc/cargo exists in this environment, so the reference itself
 * cannot be built).  Every function cites the reference file:line it follows.  Built with
 *   g++ -O2 -ffp-contract=off   (no fast-math, no FMA contraction, SSE2 doubles)
 * which is the arithmetic rustc emits for the same expressions.
 *
 * Parity pin status
 *   PINNED by the reference's own tests (ported in tests/test_kat_reference.py):
 *     Vec3 neg/add/dot/cross/length (src/utils.rs:703-772), Color::new range (:774-778),
 *     Color Display "185 200 217" (:780-785), Color inversion/add (:787-805), angle conversion
 *     (:807-831), Interval size/contains/surrounds/proportion (:833-912), Ray::at
 *     (src/camera/mod.rs:382-387), average_samples (:389-396), timeline NERP translate
 *     (src/timeline/mod.rs:329-349).
 *   PARITY UNPINNED by the reference (it has no test, fixture or golden vector for them):
 *     Sphere::hit, Triangle::hit, Aabb::hit, BVH build/traversal, all materials, textures, sky,
 *     camera ray generation, rendering.  For those this restatement IS the pin; it is defended by
 *     property tests (brute-force list vs BVH, analytic hits, degenerate-box and tie cases).
 *   The reference's RNG is an unseeded thread-local ChaCha12 (rand 0.9.2), so no bit-level pin of
 *     any random path exists even in the reference.  The oracle draws from Philox4x32-10 keyed by
 *     (seed; pixel, sample, bounce, draw) in the reference's draw ORDER.
 *   Extensions that do not exist in the reference (Quad, Emissive, black sky) are marked EXTENSION.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/crucible_gpu.h" /* POD structs of the boundary only */

namespace orc {

static const double PI = 3.14159265358979323846; /* std::f64::consts::PI */
static const double INF = std::numeric_limits<double>::infinity();

/* ------------------------------------------------------------------ utils.rs : Point3 / Vec3 */
struct V3 {
    double x, y, z;
};
/* utils.rs:248-257 */
static inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
/* utils.rs:283-293 */
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
/* utils.rs:295-301: a - b == a + (-b) */
static inline V3 sub(V3 a, V3 b) { return add(a, neg(b)); }
/* utils.rs:303-323 (f64 * Point3 and Point3 * f64 are the same products) */
static inline V3 mul(double s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
/* utils.rs:325-331: v / s == (1.0/s) * v */
static inline V3 divs(V3 v, double s) { return mul(1.0 / s, v); }
/* utils.rs:183-186: powi(2) sums left to right */
static inline double len2(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
static inline double len(V3 v) { return std::sqrt(len2(v)); } /* utils.rs:179-181 */
/* utils.rs:194-199 */
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* utils.rs:201-212 */
static inline V3 cross(V3 v, V3 o) {
    return {v.y * o.z - v.z * o.y, v.z * o.x - v.x * o.z, v.x * o.y - v.y * o.x};
}
/* utils.rs:215-218 */
static inline V3 unit(V3 v) { return divs(v, len(v)); }
/* utils.rs:189-192 */
static inline bool near_zero(V3 v) {
    const double tol = 1e-8;
    return std::fabs(v.x) < tol && std::fabs(v.y) < tol && std::fabs(v.z) < tol;
}
/* utils.rs:151-153:  v - 2.0 * v.dot(n) * n  ==  v + (-((2.0*dot) * n)) */
static inline V3 reflect(V3 v, V3 n) { return sub(v, mul(2.0 * dot(v, n), n)); }
/* Rust f64::min: returns the other operand when one is NaN */
static inline double rmin(double a, double b) { return std::fmin(a, b); }
static inline double rmax(double a, double b) { return std::fmax(a, b); }
/* Rust f64::clamp */
static inline double rclamp(double x, double lo, double hi) {
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
/* utils.rs:159-165 */
static inline V3 refract(V3 v, V3 n, double eta) {
    double cos_theta = rmin(dot(neg(v), n), 1.0);
    V3 perp = mul(eta, add(v, mul(cos_theta, n)));
    V3 par = mul(-(std::sqrt(std::fabs(1.0 - len2(perp)))), n);
    return add(perp, par);
}

/* ------------------------------------------------------------------ utils.rs : Color (clamped) */
/* A Color is an rgb triple guaranteed in [0,1] (utils.rs:340-356).  `clampc` selects the
 * reference semantics (true) or the unclamped radiance needed by the Emissive EXTENSION. */
struct Col {
    double r, g, b;
};
static inline double c01(double x, bool clampc) { return clampc ? rclamp(x, 0.0, 1.0) : x; }
/* utils.rs:445-459 Neg for Color (hilo complement) */
static inline Col col_neg(Col c) {
    double mn = (c.r < c.g ? c.r : c.g);
    mn = (mn < c.b ? mn : c.b);
    double mx = (c.r > c.g ? c.r : c.g);
    mx = (mx > c.b ? mx : c.b);
    double k = mn + mx;
    return {std::fabs(k - c.r), std::fabs(k - c.g), std::fabs(k - c.b)};
}
/* utils.rs:516-529 Add */
static inline Col col_add(Col a, Col b, bool cl) {
    return {c01(a.r + b.r, cl), c01(a.g + b.g, cl), c01(a.b + b.b, cl)};
}
/* utils.rs:558-574 f64 * Color */
static inline Col col_scale(double s, Col c, bool cl) {
    Col m = (s < 0.0) ? col_neg(c) : c;
    double p = std::fabs(s);
    return {c01(p * m.r, cl), c01(p * m.g, cl), c01(p * m.b, cl)};
}
/* utils.rs:576-590 Color * Color */
static inline Col col_mul(Col a, Col b, bool cl) {
    return {c01(a.r * b.r, cl), c01(a.g * b.g, cl), c01(a.b * b.b, cl)};
}
/* utils.rs:592-601 Color / f64 */
static inline Col col_div(Col c, double rhs, bool cl) {
    Col inv = (rhs < 0.0) ? col_neg(c) : c;
    double a = std::fabs(rhs);
    return col_scale(1.0 / a, inv, cl);
}
/* `as u32` : saturating, NaN -> 0 */
static inline uint32_t as_u32(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}
static inline int32_t as_i32(double x) {
    if (!(x == x)) return 0;
    if (x <= -2147483648.0) return INT32_MIN;
    if (x >= 2147483647.0) return INT32_MAX;
    return (int32_t)x;
}
static inline uint64_t as_usize(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)x;
}
/* utils.rs:422-438 Display for Color: byte = (255.0 * sqrt(c)) as u32 */
static inline void col_bytes(Col c, uint32_t out[3]) {
    out[0] = as_u32(255.0 * std::sqrt(c.r));
    out[1] = as_u32(255.0 * std::sqrt(c.g));
    out[2] = as_u32(255.0 * std::sqrt(c.b));
}

/* ------------------------------------------------------------------ utils.rs : Interval */
struct Interval {
    double min, max;
    double size() const { return max - min; }                                 /* :646-648 */
    bool contains(double x) const { return min <= x && x <= max; }            /* :651-653 */
    bool surrounds(double x) const { return min < x && x < max; }             /* :655-657 */
    bool is_greater(double x) const { return x < min; }                       /* :669-671 */
    bool is_less(double x) const { return x > max; }                          /* :676-678 */
    double proportion(double x) const { return (x - min) / (max - min); }     /* :682-684 */
    Interval pad(double delta) const {                                        /* :624-627 */
        double p = delta / 2.0;
        return {min - p, max + p};
    }
};
static const Interval EMPTY = {INF, -INF}; /* :695 */
/* utils.rs:631-635 */
static inline Interval tight_enclose(const Interval& a, const Interval& b) {
    return {a.min <= b.min ? a.min : b.min, a.max >= b.max ? a.max : b.max};
}

/* ------------------------------------------------------------------ camera/ray_casting.rs : Ray */
struct Ray {
    V3 o, d;
    double tm;
    V3 at(double t) const { return add(o, mul(t, d)); } /* ray_casting.rs:53-59 */
};

/* ------------------------------------------------------------------ objects/bvh.rs : Aabb */
struct Aabb {
    Interval x, y, z;
};
static const Aabb AABB_EMPTY = {EMPTY, EMPTY, EMPTY}; /* bvh.rs:25-33 */
/* bvh.rs:46-66 */
static inline Aabb aabb_from_points(V3 a, V3 b) {
    Aabb r;
    r.x = (a.x <= b.x) ? Interval{a.x, b.x} : Interval{b.x, a.x};
    r.y = (a.y <= b.y) ? Interval{a.y, b.y} : Interval{b.y, a.y};
    r.z = (a.z <= b.z) ? Interval{a.z, b.z} : Interval{b.z, a.z};
    return r;
}
/* bvh.rs:69-75 */
static inline Aabb aabb_union(const Aabb& a, const Aabb& b) {
    return {tight_enclose(a.x, b.x), tight_enclose(a.y, b.y), tight_enclose(a.z, b.z)};
}
/* bvh.rs:82-94 */
static inline int longest_axis(const Aabb& b) {
    if (b.x.size() > b.y.size()) {
        return (b.x.size() > b.z.size()) ? 0 : 2;
    } else if (b.y.size() > b.z.size()) {
        return 1;
    }
    return 2;
}
static inline const Interval& axis_interval(const Aabb& b, int ax) {
    return ax == 0 ? b.x : (ax == 1 ? b.y : b.z);
}
/* bvh.rs:96-132; ray_t is the caller's COPY (bvhwrapper.rs:98) */
static inline bool aabb_hit(const Aabb& b, const Ray& r, Interval ray_t) {
    const double o[3] = {r.o.x, r.o.y, r.o.z};
    const double d[3] = {r.d.x, r.d.y, r.d.z};
    for (int ax = 0; ax < 3; ++ax) {
        const Interval& iv = axis_interval(b, ax);
        double adinv = 1.0 / d[ax];
        double t0 = (iv.min - o[ax]) * adinv;
        double t1 = (iv.max - o[ax]) * adinv;
        double nmin, nmax;
        if (t0 < t1) {
            nmin = (t0 > ray_t.min) ? t0 : ray_t.min;
            nmax = (t1 < ray_t.max) ? t1 : ray_t.max;
        } else {
            nmin = (t1 > ray_t.min) ? t1 : ray_t.min;
            nmax = (t0 < ray_t.max) ? t0 : ray_t.max;
        }
        ray_t = {nmin, nmax};
        if (ray_t.max <= ray_t.min) return false;
    }
    return true;
}

/* ------------------------------------------------------------------ scene containers */
struct Sphere {
    V3 c;
    double r;
    int mat, obj_id, prim_index;
    bool hide;
    Aabb bbox;                   /* construction-time box: the BVH never sees update_bb (bvhwrapper.rs:47-50, 104-106) */
    std::vector<CrAnimKey> keys; /* Sphere.timeline beyond its init entries (sphere.rs:18) */
};
struct Triangle {
    V3 a, b, c;
    int mat, obj_id, prim_index;
    bool hide;
    Aabb bbox;
    std::vector<CrAnimKey> keys[3]; /* a_timeline, b_timeline, c_timeline (triangle.rs:14-16) */
};

/* TransformTimeline::combine_and_compute (timeline/mod.rs:233-263) for one point: every valid translate transform
 * (valid_time.is_less(t) || contains(t)) multiplies its matrix in, which adds its entry to that axis; of the scale
 * transforms only the LAST valid one of the list is used (`.filter(..).next_back()`, :251-257) and its matrix
 * multiplies the translated point (combined = scale * translate, :260-261; rows are dot products summed left to
 * right, the 0 * finite terms add exact zeros):
 *   kind 3 scale_sphere  diag(1,1,1,v)            (transform_builder.rs:62-80)   w = v
 *   kind 4 scale_x       diag(v,1,1,1)            (:146-164)                      x = v*x, w = 1
 *   kind 5 scale_y       row 1 = (v, 1, 0, 0)     (:229-246: v sits in the wrong slot)  y = v*x + y, w = 1
 *   kind 6 scale_z       diag(1,1,v,1)            (:312-330)                      z = v*z, w = 1
 * get_matrix_at_time (timeline/mod.rs:88-96): scaled_time = proportion(t).clamp(0, 1); a LERP scale entry is
 * start + (end - start) * scaled_time (transform_builder.rs:44-48, 128-132), a NERP one the end value.
 * Keys arrive in the timeline's sorted order. */
static inline void combine_and_compute(const std::vector<CrAnimKey>& keys, double t, V3& p, double& w) {
    int skind = -1;
    double sv = 0.0;
    for (const CrAnimKey& k : keys) {
        if (!((t > k.t1) || (k.t0 <= t && t <= k.t1))) continue;
        double s = (t - k.t0) / (k.t1 - k.t0); /* Interval::proportion, utils.rs:681-683 */
        s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s); /* f64::clamp keeps NaN */
        if (k.kind < 3) {
            double off = (k.interp == CR_LERP) ? k.a * s : k.a; /* transform_builder.rs:393-419 */
            if (k.kind == 0) p.x = off + p.x;
            else if (k.kind == 1) p.y = off + p.y;
            else p.z = off + p.z;
        } else {
            skind = k.kind;
            sv = (k.interp == CR_LERP) ? k.a + (k.b - k.a) * s : k.b;
        }
    }
    if (skind == 3) {
        w = sv;
    } else if (skind == 4) {
        p.x = sv * p.x;
        w = 1.0;
    } else if (skind == 5) {
        p.y = sv * p.x + p.y;
        w = 1.0;
    } else if (skind == 6) {
        p.z = sv * p.z;
        w = 1.0;
    }
}
struct Quad { /* EXTENSION */
    V3 q, u, v, normal, w;
    double d;
    int mat, obj_id, prim_index;
    bool hide;
    Aabb bbox;
};
struct Image {
    int w, h;
    std::vector<uint8_t> rgb;
};

enum ObjKind { O_SPHERE = 0, O_TRI = 1, O_QUAD = 2, O_NODE = 3, O_LIST = 4, O_GROUP = 5 /* a nested element before the build */ };
struct Obj {
    int kind, idx;
};
struct Node { /* bvhwrapper.rs:7-11 */
    Obj left, right;
    Aabb bbox;
};

struct HitRecord { /* objects/mod.rs:21-29 */
    V3 loc, normal;
    int mat;
    double t, u, v;
    bool front_face;
    int prim_index, obj_id;
};

struct Counters {
    uint64_t rays = 0, node = 0, sph = 0, tri = 0, quad = 0;
};

struct Scene {
    std::vector<Sphere> spheres;
    std::vector<Triangle> tris;
    std::vector<Quad> quads;
    std::vector<Obj> elements; /* every primitive in insertion order (prim_index); without groups == Scene.elements */
    /* nested elements (scene/mod.rs:160-166: add_element takes a whole HitList / BVHWrapper): the description ... */
    struct GroupDef {
        int kind; /* CR_GROUP_HITLIST / CR_GROUP_BVH */
        std::vector<Obj> members;
    };
    std::vector<GroupDef> gdefs;
    std::vector<int> open;
    std::vector<Obj> top; /* Scene.elements once a group exists: primitives and O_GROUP entries */
    /* ... and the built HitLists (hitlist.rs:6-9) */
    struct ListObj {
        std::vector<Obj> objs;
        Aabb bbox;
    };
    std::vector<ListObj> lists;
    void note_added(Obj o) {
        if (gdefs.empty()) return;
        (open.empty() ? top : gdefs[(size_t)open.back()].members).push_back(o);
    }
    std::vector<CrMaterial> mats;
    std::vector<CrTexture> texs;
    std::vector<Image> images;
    int sky_kind = CR_SKY_DEFAULT, sky_image = -1;
    bool has_emissive = false;
    /* built world */
    std::vector<Node> nodes;
    Obj world = {O_LIST, -1}; /* empty HitList */
    std::vector<uint32_t> leaf_rank; /* MODEL only: DFS rank of (leaf node, slot), for the order-free tie rule */
    struct FastTree* fast = nullptr; /* MODEL only: built on demand (orc_trace_batch mode 3) */
    bool built = false;

    const Aabb& bbox_of(const Obj& o) const {
        switch (o.kind) {
            case O_SPHERE: return spheres[o.idx].bbox;
            case O_TRI: return tris[o.idx].bbox;
            case O_QUAD: return quads[o.idx].bbox;
            case O_LIST: return o.idx < 0 ? AABB_EMPTY : lists[(size_t)o.idx].bbox; /* HitList::default(): Aabb::default() */
            default: return nodes[o.idx].bbox;
        }
    }
};

/* objects/mod.rs:38-62 HitRecord::new (normal assumed unit) */
static inline HitRecord rec_new(const Ray& r, V3 loc, V3 normal, double t, double u, double v, int mat) {
    HitRecord h;
    h.front_face = dot(r.d, normal) < 0.0;
    h.normal = h.front_face ? normal : neg(normal);
    h.loc = loc;
    h.mat = mat;
    h.t = t;
    h.u = u;
    h.v = v;
    h.prim_index = -1;
    h.obj_id = -1;
    return h;
}
/* objects/mod.rs:64-87 HitRecord::safe_new (normalises first) */
static inline HitRecord rec_safe_new(const Ray& r, V3 loc, V3 normal, double t, double u, double v, int mat) {
    return rec_new(r, loc, unit(normal), t, u, v, mat);
}

/* sphere.rs:41-46 */
static inline void sphere_uv(V3 p, double& u, double& v) {
    double theta = std::acos(-p.y);
    double phi = std::atan2(-p.z, p.x) + PI;
    u = phi / (2.0 * PI);
    v = theta / PI;
}
/* sphere.rs:61-105 (static timeline: combine_and_compute == (c, r), timeline/mod.rs:233-263) */
static inline bool sphere_hit(const Sphere& s0, const Ray& r, const Interval& ray_t, HitRecord& out) {
    if (s0.hide) return false;
    /* sphere.rs:67-70: position and radius from the timeline at the ray's time */
    struct { V3 c; double r; int mat, prim_index, obj_id; } s = {s0.c, s0.r, s0.mat, s0.prim_index, s0.obj_id};
    if (!s0.keys.empty()) combine_and_compute(s0.keys, r.tm, s.c, s.r);
    V3 oc = sub(s.c, r.o);
    double a = len2(r.d);
    double h = dot(r.d, oc);
    double c = len2(oc) - s.r * s.r;
    double disc = h * h - a * c;
    if (disc < 0.0) return false;
    double sq = std::sqrt(disc);
    double root = (h - sq) / a;
    if (!ray_t.surrounds(root)) {
        root = (h + sq) / a;
        if (!ray_t.surrounds(root)) return false;
    }
    double t = root;
    V3 p = r.at(t);
    V3 n = divs(sub(p, s.c), s.r);
    double u, v;
    sphere_uv(n, u, v);
    out = rec_new(r, p, n, t, u, v, s.mat);
    out.prim_index = s.prim_index;
    out.obj_id = s.obj_id;
    return true;
}
/* triangle.rs:86-140 */
static inline bool tri_hit(const Triangle& t0, const Ray& r, const Interval& ray_t, HitRecord& out) {
    if (t0.hide) return false;
    const double EPS = std::numeric_limits<double>::epsilon(); /* f64::EPSILON */
    /* triangle.rs:91-97: the three vertex timelines at the ray's time */
    struct { V3 a, b, c; int mat, prim_index, obj_id; } tr = {t0.a, t0.b, t0.c, t0.mat, t0.prim_index, t0.obj_id};
    double unused = 1.0;
    if (!t0.keys[0].empty()) combine_and_compute(t0.keys[0], r.tm, tr.a, unused);
    if (!t0.keys[1].empty()) combine_and_compute(t0.keys[1], r.tm, tr.b, unused);
    if (!t0.keys[2].empty()) combine_and_compute(t0.keys[2], r.tm, tr.c, unused);
    V3 e1 = sub(tr.b, tr.a);
    V3 e2 = sub(tr.c, tr.a);
    V3 pv = cross(r.d, e2);
    double det = dot(e1, pv);
    if (det > -EPS && det < EPS) return false;
    double inv = 1.0 / det;
    V3 s = sub(r.o, tr.a);
    double u = inv * dot(s, pv);
    if (!(0.0 <= u && u <= 1.0)) return false;
    V3 q = cross(s, e1);
    double v = inv * dot(r.d, q);
    if (v < 0.0 || u + v > 1.0) return false;
    double t = inv * dot(e2, q);
    if (!ray_t.surrounds(t)) return false;
    V3 p = r.at(t);
    V3 n = cross(e1, e2);
    out = rec_safe_new(r, p, n, t, 0.0, 0.0, tr.mat); /* texture u = v = 0.0, triangle.rs:133-134 */
    out.prim_index = tr.prim_index;
    out.obj_id = tr.obj_id;
    return true;
}
/* EXTENSION: quad (RTNW geometry; open interval like the reference's primitives) */
static inline bool quad_hit(const Quad& qd, const Ray& r, const Interval& ray_t, HitRecord& out) {
    if (qd.hide) return false;
    double denom = dot(qd.normal, r.d);
    if (std::fabs(denom) < 1e-8) return false;
    double t = (qd.d - dot(qd.normal, r.o)) / denom;
    if (!ray_t.surrounds(t)) return false;
    V3 p = r.at(t);
    V3 ph = sub(p, qd.q);
    double alpha = dot(qd.w, cross(ph, qd.v));
    double beta = dot(qd.w, cross(qd.u, ph));
    if (!(0.0 <= alpha && alpha <= 1.0) || !(0.0 <= beta && beta <= 1.0)) return false;
    out = rec_new(r, p, qd.normal, t, alpha, beta, qd.mat);
    out.prim_index = qd.prim_index;
    out.obj_id = qd.obj_id;
    return true;
}

static bool obj_hit(const Scene& sc, const Obj& o, const Ray& r, const Interval& ray_t, HitRecord& out, Counters& cn);

/* hitlist.rs:52-65 */
static bool list_hit(const Scene& sc, const std::vector<Obj>& objs, const Ray& r, const Interval& ray_t,
                     HitRecord& out, Counters& cn) {
    bool any = false;
    double closest = ray_t.max;
    for (const Obj& o : objs) {
        Interval iv = {ray_t.min, closest};
        HitRecord h;
        if (obj_hit(sc, o, r, iv, h, cn)) {
            closest = h.t;
            out = h;
            any = true;
        }
    }
    return any;
}
/* bvhwrapper.rs:97-126 */
static bool node_hit(const Scene& sc, const Node& n, const Ray& r, const Interval& ray_t, HitRecord& out,
                     Counters& cn) {
    cn.node++;
    if (!aabb_hit(n.bbox, r, ray_t)) return false;
    /* update_bb on both children (bvhwrapper.rs:104-106) recomputes the primitive boxes from the
     * timelines; for static timelines the result equals the construction-time box and is never
     * read by any hit test, so it is a no-op here. */
    HitRecord hl, hr;
    bool got_l = obj_hit(sc, n.left, r, ray_t, hl, cn);
    Interval rt = {ray_t.min, got_l ? hl.t : ray_t.max};
    bool got_r = obj_hit(sc, n.right, r, rt, hr, cn);
    if (got_r) {
        out = hr;
        return true;
    }
    if (got_l) {
        out = hl;
        return true;
    }
    return false;
}
/* objects/mod.rs:118-125 */
static bool obj_hit(const Scene& sc, const Obj& o, const Ray& r, const Interval& ray_t, HitRecord& out, Counters& cn) {
    switch (o.kind) {
        case O_SPHERE: cn.sph++; return sphere_hit(sc.spheres[o.idx], r, ray_t, out);
        case O_TRI: cn.tri++; return tri_hit(sc.tris[o.idx], r, ray_t, out);
        case O_QUAD: cn.quad++; return quad_hit(sc.quads[o.idx], r, ray_t, out);
        case O_NODE: return node_hit(sc, sc.nodes[o.idx], r, ray_t, out, cn);
        case O_LIST: /* nested HitList (hitlist.rs:52-65); idx -1 = the empty HitList of bvhwrapper.rs:29-31 */
            return o.idx >= 0 && list_hit(sc, sc.lists[(size_t)o.idx].objs, r, ray_t, out, cn);
        default: return false;
    }
}
static inline bool world_hit(const Scene& sc, const Ray& r, const Interval& ray_t, HitRecord& out, Counters& cn) {
    cn.rays++;
    return obj_hit(sc, sc.world, r, ray_t, out, cn);
}
/* brute force over the flat element list == HitList::hit (hitlist.rs:52-65); used by property tests */
static inline bool brute_hit(const Scene& sc, const Ray& r, const Interval& ray_t, HitRecord& out, Counters& cn) {
    return list_hit(sc, sc.elements, r, ray_t, out, cn);
}

/* ------------------------------------------------------------------ MODEL: order-free closest hit
 * NOT reference code.  The reference's closest hit (bvhwrapper.rs:97-126) has an order-free description (DESIGN.md
 * 5.1b) that the product's near-first trace kernel relies on; this is that description executed on the CPU so that
 * property tests can compare it with the reference-order traversal above, ray by ray.
 *   cand(P)   = the root Sphere::hit / Triangle::hit / quad_hit select for P on the interval (tmin, tmax): it does not
 *               depend on the running closest t, only its acceptance does;
 *   near(L)   = entry parameter of P's leaf-node box in the reference's own slab arithmetic (bvh.rs:96-132);
 *   P regular = L's box is hit on (tmin, +inf) and near(L) <= cand(P).
 * If no IRREGULAR primitive (cand(P) < near(L): rounding put the hit before its own box) has cand <= the smallest
 * regular cand, the reference returns the DFS-first regular minimiser W.  W is found in ANY visiting order; here:
 * nearer child first by the sign of the ray direction on the node's longest axis, subtrees culled when their box
 * entry exceeds best + margin.  A ray that meets an irregular candidate is flagged (the product re-traces it in
 * reference order).  cn.node counts the box tests of THIS traversal. */
struct OrderFree {
    double best;
    Obj win;
    uint32_t win_rank;
    bool has, irregular;
    double margin;
};
/* reference slab arithmetic on (tmin, +inf): returns false when the box is missed; near/far as the reference computes */
static inline bool aabb_span(const Aabb& b, const Ray& r, double tmin, double& near_t, double& far_t) {
    const double o[3] = {r.o.x, r.o.y, r.o.z};
    const double d[3] = {r.d.x, r.d.y, r.d.z};
    double lo = tmin, hi = INF;
    for (int ax = 0; ax < 3; ++ax) {
        const Interval& iv = axis_interval(b, ax);
        double adinv = 1.0 / d[ax];
        double t0 = (iv.min - o[ax]) * adinv;
        double t1 = (iv.max - o[ax]) * adinv;
        if (t0 < t1) {
            lo = (t0 > lo) ? t0 : lo;
            hi = (t1 < hi) ? t1 : hi;
        } else {
            lo = (t1 > lo) ? t1 : lo;
            hi = (t0 < hi) ? t0 : hi;
        }
        if (hi <= lo) return false;
    }
    near_t = lo; /* = max(tmin, entry) */
    far_t = hi;
    return true;
}
static void order_free_visit(const Scene& sc, const Obj& o, const Ray& r, double tmin, double tmax, OrderFree& st, Counters& cn) {
    const Node& n = sc.nodes[o.idx];
    cn.node++;
    double near_t, far_t;
    if (!aabb_span(n.bbox, r, tmin, near_t, far_t)) return;
    if (near_t > st.best + st.margin) return; /* culling only: conservative */
    if (n.left.kind == O_NODE) {
        const int ax = longest_axis(n.bbox);
        const double dax = ax == 0 ? r.d.x : (ax == 1 ? r.d.y : r.d.z);
        const bool left_first = dax >= 0.0; /* children are sorted by bbox.min on this axis (bvhwrapper.rs:66-73) */
        order_free_visit(sc, left_first ? n.left : n.right, r, tmin, tmax, st, cn);
        order_free_visit(sc, left_first ? n.right : n.left, r, tmin, tmax, st, cn);
        return;
    }
    /* leaf node: left primitive, then right (span-1 nodes hold the same primitive twice: tested once) */
    const bool same = n.left.kind == n.right.kind && n.left.idx == n.right.idx;
    for (int k = 0; k < (same ? 1 : 2); ++k) {
        const Obj& p = k == 0 ? n.left : n.right;
        HitRecord h;
        if (!obj_hit(sc, p, r, Interval{tmin, tmax}, h, cn)) continue;
        if (!(h.t < st.best || (st.has && h.t == st.best))) continue;
        /* the candidate matters: is it regular? (near_t already includes tmin <= cand) */
        if (h.t < near_t) {
            st.irregular = true;
            continue;
        }
        const uint32_t dfs = sc.leaf_rank[(size_t)2 * o.idx + k]; /* DFS position: equal cands keep the DFS-first one */
        if (h.t < st.best || dfs < st.win_rank) {
            st.best = h.t;
            st.win = p;
            st.win_rank = dfs;
            st.has = true;
        }
    }
}
/* ---- MODEL, part 2: the same order-free search over a DIFFERENT tree.  Because the result does not depend on the visiting
 * order, the search may use any conservative structure over the primitives; the reference tree is only needed for the
 * leaf-node box of a candidate (regularity) and its DFS rank (ties).  FastTree = binned-SAH binary BVH over the
 * construction-time primitive boxes, built here to measure how many box tests a good tree saves. */
struct FastNode {
    Aabb box;
    int left, right;   /* children (inner) */
    int first, count;  /* primitive range (leaf) */
};
struct FastTree {
    std::vector<FastNode> nodes;
    std::vector<Obj> prims;                       /* leaf order */
    std::vector<uint32_t> ref_leaf, ref_rank;     /* per entry of prims: reference leaf node, DFS rank */
};
static inline double half_area(const Aabb& b) {
    const double x = b.x.max - b.x.min, y = b.y.max - b.y.min, z = b.z.max - b.z.min;
    return x * y + y * z + z * x;
}
static int fast_build(const Scene& sc, FastTree& ft, std::vector<int>& idx, int lo, int hi, const std::vector<Aabb>& pb, const std::vector<V3>& pc) {
    FastNode n;
    n.box = AABB_EMPTY;
    Aabb cb = AABB_EMPTY;
    for (int i = lo; i < hi; ++i) {
        n.box = aabb_union(n.box, pb[idx[i]]);
        cb = aabb_union(cb, aabb_from_points(pc[idx[i]], pc[idx[i]]));
    }
    n.left = n.right = -1;
    n.first = lo;
    n.count = hi - lo;
    const int me = (int)ft.nodes.size();
    ft.nodes.push_back(n);
    if (hi - lo <= 2) return me;
    const int ax = longest_axis(cb);
    const Interval& ci = axis_interval(cb, ax);
    const double ext = ci.max - ci.min;
    int mid = (lo + hi) / 2;
    auto key = [&](int p) { return ax == 0 ? pc[p].x : (ax == 1 ? pc[p].y : pc[p].z); };
    if (ext > 0.0) {
        const int NB = 16;
        Aabb bb[NB];
        int bc[NB];
        for (int b = 0; b < NB; ++b) { bb[b] = AABB_EMPTY; bc[b] = 0; }
        auto bin = [&](int p) { int b = (int)((key(p) - ci.min) / ext * NB); return b < 0 ? 0 : (b >= NB ? NB - 1 : b); };
        for (int i = lo; i < hi; ++i) { int b = bin(idx[i]); bb[b] = aabb_union(bb[b], pb[idx[i]]); bc[b]++; }
        double best = 1e300; int bs = -1;
        Aabb la[NB]; int lc[NB];
        Aabb acc = AABB_EMPTY; int cnt = 0;
        for (int b = 0; b < NB; ++b) { acc = aabb_union(acc, bb[b]); cnt += bc[b]; la[b] = acc; lc[b] = cnt; }
        acc = AABB_EMPTY; cnt = 0;
        for (int b = NB - 1; b > 0; --b) {
            acc = aabb_union(acc, bb[b]); cnt += bc[b];
            if (lc[b - 1] == 0 || cnt == 0) continue;
            const double c = half_area(la[b - 1]) * lc[b - 1] + half_area(acc) * cnt;
            if (c < best) { best = c; bs = b; }
        }
        if (bs > 0) {
            int m = (int)(std::partition(idx.begin() + lo, idx.begin() + hi, [&](int p) { return bin(p) < bs; }) - idx.begin());
            if (m > lo && m < hi) mid = m;
            else std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) { return key(a) < key(b); });
        } else {
            std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) { return key(a) < key(b); });
        }
    }
    const int l = fast_build(sc, ft, idx, lo, mid, pb, pc);
    const int r = fast_build(sc, ft, idx, mid, hi, pb, pc);
    ft.nodes[me].left = l;
    ft.nodes[me].right = r;
    ft.nodes[me].count = 0;
    return me;
}
static void fast_tree_build(const Scene& sc, FastTree& ft) {
    /* primitives of the reference tree's leaf nodes, with their leaf node and DFS rank */
    std::vector<Obj> prims;
    std::vector<uint32_t> leaf, rank;
    for (size_t ni = 0; ni < sc.nodes.size(); ++ni) {
        const Node& n = sc.nodes[ni];
        if (n.left.kind == O_NODE) continue;
        const bool same = n.left.kind == n.right.kind && n.left.idx == n.right.idx;
        for (int k = 0; k < (same ? 1 : 2); ++k) {
            prims.push_back(k == 0 ? n.left : n.right);
            leaf.push_back((uint32_t)ni);
            rank.push_back(sc.leaf_rank[2 * ni + k]);
        }
    }
    std::vector<Aabb> pb(prims.size());
    std::vector<V3> pc(prims.size());
    for (size_t i = 0; i < prims.size(); ++i) {
        pb[i] = sc.bbox_of(prims[i]);
        pc[i] = {0.5 * (pb[i].x.min + pb[i].x.max), 0.5 * (pb[i].y.min + pb[i].y.max), 0.5 * (pb[i].z.min + pb[i].z.max)};
    }
    std::vector<int> idx(prims.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
    ft.nodes.clear();
    ft.nodes.reserve(prims.size());
    if (!prims.empty()) fast_build(sc, ft, idx, 0, (int)prims.size(), pb, pc);
    ft.prims.resize(prims.size());
    ft.ref_leaf.resize(prims.size());
    ft.ref_rank.resize(prims.size());
    for (size_t i = 0; i < idx.size(); ++i) {
        ft.prims[i] = prims[idx[i]];
        ft.ref_leaf[i] = leaf[idx[i]];
        ft.ref_rank[i] = rank[idx[i]];
    }
}
static void fast_visit(const Scene& sc, const FastTree& ft, int ni, const Ray& r, double tmin, double tmax, OrderFree& st, Counters& cn) {
    const FastNode& n = ft.nodes[ni];
    if (n.count > 0) {
        for (int i = n.first; i < n.first + n.count; ++i) {
            HitRecord h;
            if (!obj_hit(sc, ft.prims[i], r, Interval{tmin, tmax}, h, cn)) continue;
            if (!(h.t < st.best || (st.has && h.t == st.best))) continue;
            double near_t, far_t; /* the candidate matters: consult the REFERENCE leaf-node box */
            if (!aabb_span(sc.nodes[ft.ref_leaf[i]].bbox, r, tmin, near_t, far_t)) continue; /* the reference never tests it */
            if (h.t < near_t) { st.irregular = true; continue; }
            if (h.t < st.best || ft.ref_rank[i] < st.win_rank) {
                st.best = h.t; st.win = ft.prims[i]; st.win_rank = ft.ref_rank[i]; st.has = true;
            }
        }
        return;
    }
    double nl, fl, nr, fr;
    cn.node += 2;
    bool hl = aabb_span(ft.nodes[n.left].box, r, tmin, nl, fl) && nl <= st.best + st.margin;
    bool hr = aabb_span(ft.nodes[n.right].box, r, tmin, nr, fr) && nr <= st.best + st.margin;
    if (hl && hr) {
        const bool lf = nl <= nr;
        fast_visit(sc, ft, lf ? n.left : n.right, r, tmin, tmax, st, cn);
        const double nn = lf ? nr : nl;
        if (nn <= st.best + st.margin) fast_visit(sc, ft, lf ? n.right : n.left, r, tmin, tmax, st, cn);
    } else if (hl) fast_visit(sc, ft, n.left, r, tmin, tmax, st, cn);
    else if (hr) fast_visit(sc, ft, n.right, r, tmin, tmax, st, cn);
}

static void assign_leaf_ranks(Scene& sc, const Obj& o, uint32_t& next) {
    if (o.kind != O_NODE) return;
    const Node& n = sc.nodes[o.idx];
    if (n.left.kind == O_NODE) {
        assign_leaf_ranks(sc, n.left, next);
        assign_leaf_ranks(sc, n.right, next);
    } else {
        sc.leaf_rank[(size_t)2 * o.idx] = next++;
        sc.leaf_rank[(size_t)2 * o.idx + 1] = next++;
    }
}
/* returns: 1 hit, 0 miss, -1 flagged irregular (caller falls back to world_hit) */
static inline int order_free_hit(const Scene& sc, const Ray& r, double tmin, double tmax, double margin_k, HitRecord& out, Counters& cn, const FastTree* ft = nullptr) {
    cn.rays++;
    if (sc.world.kind != O_NODE) return 0;
    const double dl = std::sqrt(len2(r.d));
    const double oo = std::fabs(r.o.x) + std::fabs(r.o.y) + std::fabs(r.o.z);
    const Aabb& rb = sc.nodes[sc.world.idx].bbox;
    double B = 0.0;
    for (int ax = 0; ax < 3; ++ax) {
        const Interval& iv = axis_interval(rb, ax);
        B = std::max(B, std::max(std::fabs(iv.min), std::fabs(iv.max)));
    }
    OrderFree st;
    st.best = tmax;
    st.win = {O_LIST, -1};
    st.win_rank = 0xFFFFFFFFu;
    st.has = false;
    st.irregular = false;
    st.margin = margin_k * (oo + 3.0 * B) / dl;
    if (ft) {
        double n0, f0;
        cn.node++;
        if (!ft->nodes.empty() && aabb_span(ft->nodes[0].box, r, tmin, n0, f0)) fast_visit(sc, *ft, 0, r, tmin, tmax, st, cn);
    } else {
        order_free_visit(sc, sc.world, r, tmin, tmax, st, cn);
    }
    if (st.irregular) return -1;
    if (!st.has) return 0;
    Counters dummy;
    obj_hit(sc, st.win, r, Interval{tmin, tmax}, out, dummy);
    return 1;
}

/* ------------------------------------------------------------------ bvhwrapper.rs : build */
/* bvhwrapper.rs:82-94 */
static inline bool box_less(const Scene& sc, const Obj& a, const Obj& b, int axis) {
    return axis_interval(sc.bbox_of(a), axis).min < axis_interval(sc.bbox_of(b), axis).min;
}
/* bvhwrapper.rs:46-80 */
static Obj help_generate(Scene& sc, std::vector<Obj>& objects, size_t start, size_t end) {
    Aabb bbox = AABB_EMPTY;
    for (size_t i = start; i < end; ++i) bbox = aabb_union(bbox, sc.bbox_of(objects[i]));
    int axis = longest_axis(bbox);
    size_t span = end - start;
    Obj left, right;
    if (span == 1) {
        left = objects[start];
        right = objects[start];
    } else if (span == 2) {
        left = objects[start];
        right = objects[start + 1];
    } else {
        std::stable_sort(objects.begin() + start, objects.begin() + end,
                         [&](const Obj& a, const Obj& b) { return box_less(sc, a, b, axis); });
        size_t mid = start + span / 2;
        left = help_generate(sc, objects, start, mid);
        right = help_generate(sc, objects, mid, end);
    }
    Node n;
    n.left = left;
    n.right = right;
    n.bbox = bbox;
    sc.nodes.push_back(n);
    return {O_NODE, (int)sc.nodes.size() - 1};
}
static Obj resolve_element(Scene& sc, const Obj& o);
/* BVHWrapper::new_wrapper, bvhwrapper.rs:15-44: hidden primitives dropped, nested lists / wrappers kept */
static Obj new_wrapper(Scene& sc, const std::vector<Obj>& objs) {
    std::vector<Obj> visible;
    for (const Obj& o : objs) {
        bool hide = (o.kind == O_SPHERE) ? sc.spheres[o.idx].hide
                    : (o.kind == O_TRI)  ? sc.tris[o.idx].hide
                    : (o.kind == O_QUAD) ? sc.quads[o.idx].hide
                                         : false;
        if (!hide) visible.push_back(resolve_element(sc, o));
    }
    if (visible.empty()) return {O_LIST, -1};
    Obj root = help_generate(sc, visible, 0, visible.size());
    const Aabb fixed = aabb_union(sc.bbox_of(sc.nodes[root.idx].left), sc.bbox_of(sc.nodes[root.idx].right)); /* new_from_vec, :38-41 */
    sc.nodes[root.idx].bbox = fixed;
    return root;
}
/* a nested element as the caller built it before Scene::add_element: HitList::add per member (hitlist.rs:27-30), or
 * BVHWrapper::new_wrapper over the members */
static Obj resolve_element(Scene& sc, const Obj& o) {
    if (o.kind != O_GROUP) return o;
    const int kind = sc.gdefs[(size_t)o.idx].kind;
    const std::vector<Obj> members = sc.gdefs[(size_t)o.idx].members;
    if (kind == CR_GROUP_BVH) return new_wrapper(sc, members);
    Scene::ListObj l;
    l.bbox = AABB_EMPTY;
    for (const Obj& m : members) {
        const Obj r = resolve_element(sc, m);
        l.objs.push_back(r);
        l.bbox = aabb_union(l.bbox, sc.bbox_of(r));
    }
    sc.lists.push_back(l);
    return {O_LIST, (int)sc.lists.size() - 1};
}
/* Scene::render_image: BVHWrapper::new_wrapper(self.elements.clone()), scene/mod.rs:333 */
static void build_world(Scene& sc) {
    sc.nodes.clear();
    sc.lists.clear();
    sc.world = new_wrapper(sc, sc.gdefs.empty() ? sc.elements : sc.top);
    delete sc.fast; /* MODEL tree of the previous build */
    sc.fast = nullptr;
    sc.leaf_rank.assign(2 * sc.nodes.size(), 0u);
    {
        uint32_t next = 0;
        if (sc.gdefs.empty()) assign_leaf_ranks(sc, sc.world, next); /* the order-free MODEL covers flat scenes only */
    }
    sc.built = true;
}

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al. 2011) */
static inline void philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* One stream per (pixel, sample, bounce); draw k comes from block k/2, half k%2:
 * u = ((hi<<32 | lo) >> 11) * 2^-53 in [0,1).  bounce 0 = camera sample, bounce b>=1 = b-th scatter. */
struct Rng {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int have = 0;
    Rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = pixel; ctr[1] = sample; ctr[2] = bounce; ctr[3] = 0;
    }
    double next() {
        if (have == 0) {
            philox(ctr, key, buf);
            ctr[3]++;
            have = 2;
        }
        int h = 2 - have;
        have--;
        uint64_t x = ((uint64_t)buf[2 * h + 1] << 32) | buf[2 * h];
        return (double)(x >> 11) * (1.0 / 9007199254740992.0);
    }
    double range(double lo, double hi) { return lo + (hi - lo) * next(); }
};
/* utils.rs:128-138 (draw order of random_vec3_range: x, y, z) */
static inline V3 random_unit_vector(Rng& g) {
    for (;;) {
        double x = g.range(-1.0, 1.0);
        double y = g.range(-1.0, 1.0);
        double z = g.range(-1.0, 1.0);
        V3 p = {x, y, z};
        double l2 = len2(p);
        if (1e-160 < l2 && l2 <= 1.0) return divs(p, std::sqrt(l2));
    }
}
/* utils.rs:112-126 */
static inline V3 random_in_unit_disk(Rng& g) {
    for (;;) {
        double x = g.range(-1.0, 1.0);
        double y = g.range(-1.0, 1.0);
        V3 p = {x, y, 0.0};
        if (len2(p) < 1.0) return p;
    }
}

/* ------------------------------------------------------------------ textures */
/* asset_loader/img_loader.rs:69-77 + :40-42 */
static inline Col image_pixel(const Image& im, uint64_t x, uint64_t y) {
    uint64_t xm = (uint64_t)im.w - 1, ym = (uint64_t)im.h - 1;
    if (x > xm) x = xm;
    if (y > ym) y = ym;
    const uint8_t* p = &im.rgb[(y * (uint64_t)im.w + x) * 3];
    return {p[0] / 255.0, p[1] / 255.0, p[2] / 255.0};
}
/* textures/image_texture.rs:23-32 and scene/mod.rs:37-45 share this arithmetic */
static inline Col image_lookup(const Image& im, double u, double v) {
    u = rclamp(u, 0.0, 1.0);
    v = 1.0 - rclamp(v, 0.0, 1.0);
    uint64_t i = as_usize(u * (double)im.w);
    uint64_t j = as_usize(v * (double)im.h);
    return image_pixel(im, i, j);
}
/* textures/mod.rs:20-26 */
static Col tex_value(const Scene& sc, int tex, double u, double v, V3 p) {
    const CrTexture& t = sc.texs[tex];
    switch (t.kind) {
        case CR_TEX_SOLID: return {t.color[0], t.color[1], t.color[2]}; /* solid_color.rs:25-27 */
        case CR_TEX_CHECKER: {                                          /* checker_texture.rs:39-51 */
            int32_t xi = as_i32(std::floor(t.inv_scale * p.x));
            int32_t yi = as_i32(std::floor(t.inv_scale * p.y));
            int32_t zi = as_i32(std::floor(t.inv_scale * p.z));
            int32_t sum = (int32_t)((uint32_t)xi + (uint32_t)yi + (uint32_t)zi);
            bool even = (sum % 2) == 0;
            return tex_value(sc, even ? t.even : t.odd, u, v, p);
        }
        default: return image_lookup(sc.images[t.image], u, v); /* image_texture.rs:23-32 */
    }
}

/* ------------------------------------------------------------------ materials */
/* returns true when the ray scatters; att is always written (reference writes it before deciding) */
static bool scatter(const Scene& sc, const Ray& r_in, const HitRecord& rec, Col& att, Ray& out, Rng& g, bool cl) {
    const CrMaterial& m = sc.mats[rec.mat];
    switch (m.kind) {
        case CR_MAT_LAMBERTIAN: { /* lambertian.rs:40-61 */
            V3 dir = add(rec.normal, random_unit_vector(g));
            if (near_zero(dir)) dir = rec.normal;
            out = {rec.loc, dir, r_in.tm};
            att = col_div(tex_value(sc, m.tex, rec.u, rec.v, rec.loc), m.scatter_prob, true);
            return g.next() <= m.scatter_prob;
        }
        case CR_MAT_METAL: { /* metal.rs:29-42 */
            V3 refl = reflect(r_in.d, rec.normal);
            V3 dir = add(unit(refl), mul(m.fuzz, random_unit_vector(g)));
            out = {rec.loc, dir, r_in.tm};
            att = {m.albedo[0], m.albedo[1], m.albedo[2]};
            return dot(out.d, rec.normal) > 0.0;
        }
        case CR_MAT_DIELECTRIC: { /* dielectric.rs:30-55 */
            att = {1.0, 1.0, 1.0};
            double ri = rec.front_face ? 1.0 / m.ior : m.ior;
            V3 ud = unit(r_in.d);
            double cos_theta = -(rmin(dot(ud, rec.normal), 1.0)); /* precedence of :40 */
            double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
            bool cannot = ri * sin_theta > 1.0;
            bool refl = cannot;
            if (!refl) { /* short-circuit: RNG drawn only when refraction is possible */
                double r0 = (1.0 - ri) / (1.0 + ri);
                r0 = r0 * r0;
                double x = 1.0 - cos_theta;
                double x2 = x * x;
                double x5 = x * (x2 * x2); /* powi(5), SURVEY App. B */
                double reflectance = r0 + (1.0 - r0) * x5;
                refl = reflectance > g.next();
            }
            V3 dir = refl ? reflect(ud, rec.normal) : refract(ud, rec.normal, ri);
            out = {rec.loc, dir, r_in.tm};
            return true;
        }
        default: /* EXTENSION Emissive: never scatters */
            (void)cl;
            att = {0.0, 0.0, 0.0};
            return false;
    }
}

/* ------------------------------------------------------------------ sky (ray_casting.rs:133-151) */
static Col sky_color(const Scene& sc, const Ray& r, bool cl) {
    if (sc.sky_kind == CR_SKY_SPHERICAL) {
        V3 ud = unit(r.d);
        double theta = std::atan2(ud.x, ud.z);
        double phi = std::asin(ud.y);
        double u = (theta / (2.0 * PI)) + 0.5;
        double v = (phi / PI) + 0.5;
        return image_lookup(sc.images[sc.sky_image], u, v);
    }
    if (sc.sky_kind == CR_SKY_BLACK) return {0.0, 0.0, 0.0}; /* EXTENSION */
    V3 ud = unit(r.d);
    double a = 0.5 * (ud.y + 1.0);
    return col_add(col_scale(1.0 - a, Col{1.0, 1.0, 1.0}, cl), col_scale(a, Col{0.5, 0.7, 1.0}, cl), cl);
}

/* ray_casting.rs:112-152 (recursive, as in the reference).  bounce = index of the NEXT scatter. */
static Col ray_color(const Scene& sc, const Ray& r, uint32_t depth, uint64_t seed, uint32_t pixel, uint32_t sample,
                     uint32_t bounce, Counters& cn) {
    const bool cl = !sc.has_emissive;
    if (depth == 0) return {0.0, 0.0, 0.0};
    HitRecord h;
    if (world_hit(sc, r, Interval{0.001, INF}, h, cn)) {
        Col att = {0.0, 0.0, 0.0};
        Ray s;
        Rng g(seed, pixel, sample, bounce);
        const CrMaterial& m = sc.mats[h.mat];
        if (m.kind == CR_MAT_EMISSIVE) return {m.emit[0], m.emit[1], m.emit[2]}; /* EXTENSION */
        if (scatter(sc, r, h, att, s, g, cl)) {
            return col_mul(att, ray_color(sc, s, depth - 1, seed, pixel, sample, bounce + 1, cn), cl);
        }
        return {0.0, 0.0, 0.0};
    }
    return sky_color(sc, r, cl);
}

/* ------------------------------------------------------------------ camera */
/* timeline/mod.rs:233-263 for a point with translate keyframes only: the product of translation
 * matrices adds the offsets in timeline order; scale is the identity (start_scale 1.0). */
static V3 point_at(const double init[3], const CrKeyframe* keys, uint32_t n, double t) {
    double p[3] = {init[0], init[1], init[2]};
    for (uint32_t k = 0; k < n; ++k) {
        Interval iv = {keys[k].t0, keys[k].t1};
        if (!(iv.is_less(t) || iv.contains(t))) continue; /* :239-243 */
        double s = rclamp(iv.proportion(t), 0.0, 1.0);    /* :92 */
        double off = (keys[k].interp == CR_LERP) ? keys[k].delta * s : keys[k].delta;
        p[keys[k].axis] = off + p[keys[k].axis];
    }
    return {p[0], p[1], p[2]};
}
struct Cam {
    CrCamera c;
    V3 get_from(double t) const { return point_at(c.look_from, c.from_keys, c.n_from_keys, t); } /* camera/mod.rs:319-322 */
    V3 get_at(double t) const { return point_at(c.look_at, c.at_keys, c.n_at_keys, t); }         /* :324-327 */
    V3 vup() const { return {c.vup[0], c.vup[1], c.vup[2]}; }
    /* rendering_compute.rs:88-93 */
    V3 w_basis(double t) const { return unit(sub(get_from(t), get_at(t))); }
    /* :77-80 */
    V3 u_basis(double t) const { return unit(cross(vup(), w_basis(t))); }
    /* :82-85 */
    V3 v_basis(double t) const { return cross(w_basis(t), u_basis(t)); }
    /* :16-19 */
    V3 viewport_u(double t) const { return mul(c.viewport_width, u_basis(t)); }
    /* :24-27 */
    V3 viewport_v(double t) const { return mul(c.viewport_height, neg(v_basis(t))); }
    /* :32-35, :40-43 */
    V3 pixel_delta_u(double t) const { return divs(viewport_u(t), (double)c.image_width); }
    V3 pixel_delta_v(double t) const { return divs(viewport_v(t), (double)c.image_height); }
    /* :49-55 */
    V3 viewport_upperleft(double t) const {
        V3 cc = get_from(t);
        return sub(sub(sub(cc, mul(c.focus_dist, w_basis(t))), divs(viewport_u(t), 2.0)), divs(viewport_v(t), 2.0));
    }
    /* :57-60 */
    V3 pixel_start_location(double t) const {
        return add(viewport_upperleft(t), mul(0.5, add(pixel_delta_u(t), pixel_delta_v(t))));
    }
    /* :64-68 */
    V3 get_pixel_pos(uint32_t i, uint32_t j, V3 off, double t) const {
        return add(add(pixel_start_location(t), mul((double)i + off.x, pixel_delta_u(t))),
                   mul((double)j + off.y, pixel_delta_v(t)));
    }
    /* :95-103: Vec3 * f64 multiplies rhs * component */
    V3 defocus_disk_u(double t) const { return mul(c.defocus_radius, u_basis(t)); }
    V3 defocus_disk_v(double t) const { return mul(c.defocus_radius, v_basis(t)); }
    /* :105-110 */
    V3 defocus_disk_sample(double t, Rng& g) const {
        V3 p = random_in_unit_disk(g);
        V3 from = get_from(t);
        return add(add(from, mul(p.x, defocus_disk_u(t))), mul(p.y, defocus_disk_v(t)));
    }
    double current_time() const { return (double)c.frame * (1.0 / c.frame_rate); }              /* ray_casting.rs:77 */
    double shutter_length() const { return (c.shutter_angle / 360.0) * (1.0 / c.frame_rate); }   /* :79 */
    /* one iteration of the sample loop, ray_casting.rs:82-104 */
    Ray sample_ray(uint32_t i, uint32_t j, Rng& g) const {
        double time_sample = current_time() + g.range(0.0, shutter_length());
        V3 cc = get_from(time_sample);
        double ox = g.next() - 0.5; /* sample_square, camera/mod.rs:369-376 */
        double oy = g.next() - 0.5;
        V3 ps = get_pixel_pos(i, j, V3{ox, oy, 0.0}, time_sample);
        V3 orig = (c.defocus_angle <= 0.0) ? cc : defocus_disk_sample(time_sample, g);
        return {orig, sub(ps, orig), time_sample};
    }
};

/* ray_casting.rs:64-108 + average_samples :154-173 */
static Col cast_ray(const Scene& sc, const Cam& cam, uint32_t i, uint32_t j, uint64_t seed, Counters& cn) {
    uint32_t pixel = j * cam.c.image_width + i;
    double rt = 0.0, gt = 0.0, bt = 0.0;
    for (uint32_t s = 0; s < cam.c.samples; ++s) {
        Rng g(seed, pixel, s, 0);
        Ray r = cam.sample_ray(i, j, g);
        Col c = ray_color(sc, r, cam.c.max_depth, seed, pixel, s, 1, cn);
        rt += c.r;
        gt += c.g;
        bt += c.b;
    }
    double n = (double)cam.c.samples;
    rt /= n;
    gt /= n;
    bt /= n;
    if (sc.has_emissive) { /* EXTENSION: radiance can exceed 1; the reference's Color::new would panic */
        rt = rclamp(rt, 0.0, 1.0);
        gt = rclamp(gt, 0.0, 1.0);
        bt = rclamp(bt, 0.0, 1.0);
    }
    return {rt, gt, bt};
}

static thread_local std::string g_err;

} /* namespace orc */

using namespace orc;

extern "C" {

typedef struct OrcScene OrcScene;

const char* orc_last_error(void) { return g_err.c_str(); }

OrcScene* orc_scene_create(void) { return reinterpret_cast<OrcScene*>(new Scene()); }
void orc_scene_destroy(OrcScene* s) { delete reinterpret_cast<Scene*>(s); }

int64_t orc_scene_add_spheres(OrcScene* h, const double* d, const int32_t* mat, const int32_t* obj_id, size_t n) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    int64_t first = (int64_t)sc.elements.size();
    for (size_t i = 0; i < n; ++i) {
        Sphere s;
        s.c = {d[4 * i], d[4 * i + 1], d[4 * i + 2]};
        s.r = d[4 * i + 3];
        s.mat = mat ? mat[i] : 0;
        s.obj_id = obj_id ? obj_id[i] : (int)sc.elements.size();
        s.prim_index = (int)sc.elements.size();
        s.hide = false;
        V3 rv = {s.r, s.r, s.r};
        s.bbox = aabb_from_points(sub(s.c, rv), add(s.c, rv)); /* sphere.rs:29-30 */
        sc.spheres.push_back(s);
        sc.elements.push_back({O_SPHERE, (int)sc.spheres.size() - 1});
        sc.note_added(sc.elements.back());
    }
    sc.built = false;
    return first;
}
int64_t orc_scene_add_triangles(OrcScene* h, const double* d, const int32_t* mat, const int32_t* obj_id, size_t n) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    int64_t first = (int64_t)sc.elements.size();
    for (size_t i = 0; i < n; ++i) {
        Triangle t;
        const double* p = d + 9 * i;
        t.a = {p[0], p[1], p[2]};
        t.b = {p[3], p[4], p[5]};
        t.c = {p[6], p[7], p[8]};
        t.mat = mat ? mat[i] : 0;
        t.obj_id = obj_id ? obj_id[i] : (int)sc.elements.size();
        t.prim_index = (int)sc.elements.size();
        t.hide = false;
        /* triangle.rs:28-36, 48-62: a.max(b.max(c)) / a.min(b.min(c)) per axis */
        t.bbox.x = {rmin(t.a.x, rmin(t.b.x, t.c.x)), rmax(t.a.x, rmax(t.b.x, t.c.x))};
        t.bbox.y = {rmin(t.a.y, rmin(t.b.y, t.c.y)), rmax(t.a.y, rmax(t.b.y, t.c.y))};
        t.bbox.z = {rmin(t.a.z, rmin(t.b.z, t.c.z)), rmax(t.a.z, rmax(t.b.z, t.c.z))};
        sc.tris.push_back(t);
        sc.elements.push_back({O_TRI, (int)sc.tris.size() - 1});
        sc.note_added(sc.elements.back());
    }
    sc.built = false;
    return first;
}
/* EXTENSION */
int64_t orc_scene_add_quads(OrcScene* h, const double* d, const int32_t* mat, const int32_t* obj_id, size_t n) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    int64_t first = (int64_t)sc.elements.size();
    for (size_t i = 0; i < n; ++i) {
        Quad q;
        const double* p = d + 9 * i;
        q.q = {p[0], p[1], p[2]};
        q.u = {p[3], p[4], p[5]};
        q.v = {p[6], p[7], p[8]};
        V3 nn = cross(q.u, q.v);
        q.normal = unit(nn);
        q.d = dot(q.normal, q.q);
        q.w = divs(nn, dot(nn, nn));
        q.mat = mat ? mat[i] : 0;
        q.obj_id = obj_id ? obj_id[i] : (int)sc.elements.size();
        q.prim_index = (int)sc.elements.size();
        q.hide = false;
        Aabb d1 = aabb_from_points(q.q, add(add(q.q, q.u), q.v));
        Aabb d2 = aabb_from_points(add(q.q, q.u), add(q.q, q.v));
        Aabb bb = aabb_union(d1, d2);
        const double delta = 0.0001;
        if (bb.x.size() < delta) bb.x = bb.x.pad(delta);
        if (bb.y.size() < delta) bb.y = bb.y.pad(delta);
        if (bb.z.size() < delta) bb.z = bb.z.pad(delta);
        q.bbox = bb;
        sc.quads.push_back(q);
        sc.elements.push_back({O_QUAD, (int)sc.quads.size() - 1});
        sc.note_added(sc.elements.back());
    }
    sc.built = false;
    return first;
}
/* nested elements (mirror of cr_scene_begin_group / cr_scene_end_group) */
int orc_scene_begin_group(OrcScene* h, int kind) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    if (kind != CR_GROUP_HITLIST && kind != CR_GROUP_BVH) return CR_ERR_INVALID;
    if (sc.gdefs.empty()) sc.top = sc.elements; /* everything added so far is a top-level element */
    const int id = (int)sc.gdefs.size();
    Scene::GroupDef g;
    g.kind = kind;
    (sc.open.empty() ? sc.top : sc.gdefs[(size_t)sc.open.back()].members).push_back({O_GROUP, id});
    sc.gdefs.push_back(g);
    sc.open.push_back(id);
    sc.built = false;
    return id;
}
int orc_scene_end_group(OrcScene* h) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    if (sc.open.empty()) return CR_ERR_STATE;
    sc.open.pop_back();
    return 0;
}
/* keyframes of one point of a primitive, in the timeline's sorted order (mirror of cr_scene_set_keyframes) */
int orc_scene_set_keyframes(OrcScene* h, size_t prim, int point, const CrAnimKey* keys, size_t n) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    if (prim >= sc.elements.size()) return CR_ERR_INVALID;
    Obj o = sc.elements[prim];
    if (o.kind == O_SPHERE && point == 0) sc.spheres[o.idx].keys.assign(keys, keys + n);
    else if (o.kind == O_TRI && point >= 0 && point <= 2) sc.tris[o.idx].keys[point].assign(keys, keys + n);
    else return CR_ERR_INVALID;
    return 0;
}
/* TransformTimeline::combine_and_compute on its own (timeline KATs): out = (x, y, z, w) */
void orc_combine_and_compute(const double init[4], const CrAnimKey* keys, size_t n, double t, double out[4]) {
    std::vector<CrAnimKey> v(keys, keys + n);
    V3 p = {init[0], init[1], init[2]};
    double w = init[3];
    combine_and_compute(v, t, p, w);
    out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = w;
}
int orc_scene_set_hidden(OrcScene* h, size_t prim, int hide) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    if (prim >= sc.elements.size()) return CR_ERR_INVALID;
    Obj o = sc.elements[prim];
    if (o.kind == O_SPHERE) sc.spheres[o.idx].hide = hide != 0;
    else if (o.kind == O_TRI) sc.tris[o.idx].hide = hide != 0;
    else sc.quads[o.idx].hide = hide != 0;
    sc.built = false;
    return 0;
}
int orc_scene_set_materials(OrcScene* h, const CrMaterial* m, size_t n) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    sc.mats.assign(m, m + n);
    sc.has_emissive = false;
    for (auto& x : sc.mats) sc.has_emissive |= (x.kind == CR_MAT_EMISSIVE);
    return 0;
}
int orc_scene_set_textures(OrcScene* h, const CrTexture* t, size_t n) {
    reinterpret_cast<Scene*>(h)->texs.assign(t, t + n);
    return 0;
}
int orc_scene_add_image(OrcScene* h, const uint8_t* rgb, int w, int hgt) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    Image im;
    im.w = w;
    im.h = hgt;
    im.rgb.assign(rgb, rgb + (size_t)w * hgt * 3);
    sc.images.push_back(std::move(im));
    return (int)sc.images.size() - 1;
}
int orc_scene_set_sky(OrcScene* h, int kind, int image) {
    Scene& sc = *reinterpret_cast<Scene*>(h);
    sc.sky_kind = kind;
    sc.sky_image = image;
    return 0;
}
int orc_scene_commit(OrcScene* h) {
    build_world(*reinterpret_cast<Scene*>(h));
    return 0;
}
/* DFS leaf order of the built tree (left subtree first), as prim_index values */
static void leaf_order(const Scene& sc, const Obj& o, std::vector<int32_t>& out, uint32_t depth, uint32_t& maxd) {
    if (depth > maxd) maxd = depth;
    switch (o.kind) {
        case O_SPHERE: out.push_back(sc.spheres[o.idx].prim_index); break;
        case O_TRI: out.push_back(sc.tris[o.idx].prim_index); break;
        case O_QUAD: out.push_back(sc.quads[o.idx].prim_index); break;
        case O_NODE:
            leaf_order(sc, sc.nodes[o.idx].left, out, depth + 1, maxd);
            leaf_order(sc, sc.nodes[o.idx].right, out, depth + 1, maxd);
            break;
        case O_LIST:
            if (o.idx >= 0)
                for (const Obj& m : sc.lists[(size_t)o.idx].objs) leaf_order(sc, m, out, depth + 1, maxd);
            break;
        default: break;
    }
}
int orc_scene_bvh_info(const OrcScene* h, uint64_t* n_nodes, uint32_t* max_depth, uint64_t* n_visible) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    std::vector<int32_t> order;
    uint32_t maxd = 0;
    leaf_order(sc, sc.world, order, 0, maxd);
    if (n_nodes) *n_nodes = sc.nodes.size();
    if (max_depth) *max_depth = maxd;
    if (n_visible) {
        /* span-1 nodes list their primitive twice */
        std::vector<int32_t> u = order;
        std::sort(u.begin(), u.end());
        *n_visible = (uint64_t)(std::unique(u.begin(), u.end()) - u.begin());
    }
    return 0;
}
int64_t orc_scene_bvh_leaf_order(const OrcScene* h, int32_t* out, size_t cap) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    std::vector<int32_t> order;
    uint32_t maxd = 0;
    leaf_order(sc, sc.world, order, 0, maxd);
    size_t n = std::min(cap, order.size());
    if (out) memcpy(out, order.data(), n * sizeof(int32_t));
    return (int64_t)order.size();
}
/* Root box (tests of the build) */
int orc_scene_root_bbox(const OrcScene* h, double out[6]) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    if (sc.world.kind != O_NODE) return CR_ERR_STATE;
    const Aabb& b = sc.nodes[sc.world.idx].bbox;
    out[0] = b.x.min; out[1] = b.x.max; out[2] = b.y.min; out[3] = b.y.max; out[4] = b.z.min; out[5] = b.z.max;
    return 0;
}

static void fill_hit(const HitRecord& h, bool got, CrHit& o) {
    if (!got) {
        memset(&o, 0, sizeof(o));
        o.prim_index = -1;
        o.obj_id = -1;
        o.material = -1;
        return;
    }
    o.prim_index = h.prim_index;
    o.obj_id = h.obj_id;
    o.front_face = h.front_face ? 1 : 0;
    o.material = h.mat;
    o.t = h.t;
    o.p[0] = h.loc.x; o.p[1] = h.loc.y; o.p[2] = h.loc.z;
    o.n[0] = h.normal.x; o.n[1] = h.normal.y; o.n[2] = h.normal.z;
    o.u = h.u;
    o.v = h.v;
}

static double g_order_free_margin = 9.5367431640625e-07; /* 2^-20, the product's margin factor */
extern "C" void orc_set_order_free_margin(double k) { g_order_free_margin = k; }
/* Hittables::hit on a ray batch.  mode 0 = built world (BVH), 1 = brute-force flat list, 2 = order-free MODEL.
 * counters (optional) = [n][4] u32: aabb tests, sphere tests, triangle tests, quad tests. */
int orc_trace_batch(const OrcScene* h, const double* rays, size_t n, double tmin, double tmax, int mode, CrHit* out,
                    uint32_t* counters, int nthreads) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    if (!sc.built) {
        g_err = "orc_trace_batch: scene not committed";
        return CR_ERR_STATE;
    }
    if (nthreads < 1) nthreads = 1;
    if (mode == 3 && !sc.fast) {
        Scene& msc = const_cast<Scene&>(sc);
        msc.fast = new FastTree();
        fast_tree_build(sc, *msc.fast);
    }
    std::atomic<size_t> next{0};
    auto work = [&]() {
        const size_t chunk = 4096;
        for (;;) {
            size_t b = next.fetch_add(chunk);
            if (b >= n) break;
            size_t e = std::min(n, b + chunk);
            for (size_t i = b; i < e; ++i) {
                const double* p = rays + 7 * i;
                Ray r = {{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, p[6]};
                HitRecord hr;
                Counters cn;
                bool got;
                if (mode >= 2) { /* MODEL of the order-free traversal; counters[3] = 0xFFFFFFFF marks a flagged ray */
                    const int rc = order_free_hit(sc, r, tmin, tmax, g_order_free_margin, hr, cn, mode == 3 ? sc.fast : nullptr);
                    got = rc > 0;
                    if (rc < 0) {
                        Counters c2;
                        got = world_hit(sc, r, Interval{tmin, tmax}, hr, c2);
                        cn.quad = 0xFFFFFFFFu;
                    }
                } else {
                    got = (mode == 0) ? world_hit(sc, r, Interval{tmin, tmax}, hr, cn)
                                      : brute_hit(sc, r, Interval{tmin, tmax}, hr, cn);
                }
                fill_hit(hr, got, out[i]);
                if (counters) {
                    counters[4 * i + 0] = (uint32_t)cn.node;
                    counters[4 * i + 1] = (uint32_t)cn.sph;
                    counters[4 * i + 2] = (uint32_t)cn.tri;
                    counters[4 * i + 3] = (uint32_t)cn.quad;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return 0;
}

typedef struct OrcRenderStats {
    uint64_t samples, rays, node_tests, sphere_tests, tri_tests, quad_tests;
    double seconds;
    int32_t threads;
    int32_t pad;
} OrcRenderStats;

/* Camera::render sample loop (camera/mod.rs:284-303) over rows [row_begin,row_end) step row_step,
 * multithreaded over rows with the world shared read-only (BASELINE.md CPU-baseline plan).
 * out_rgb [H][W][3] f64 (rows not rendered are left untouched), out_rgb8 optional. */
int orc_render(const OrcScene* h, const CrCamera* cam_in, uint64_t seed, uint32_t row_begin, uint32_t row_end,
               uint32_t row_step, int nthreads, double* out_rgb, uint8_t* out_rgb8, OrcRenderStats* st) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    if (!sc.built) {
        g_err = "orc_render: scene not committed";
        return CR_ERR_STATE;
    }
    Cam cam;
    cam.c = *cam_in;
    if (row_end > cam.c.image_height) row_end = cam.c.image_height;
    if (row_step == 0) row_step = 1;
    if (nthreads < 1) nthreads = 1;
    std::vector<uint32_t> rows;
    for (uint32_t j = row_begin; j < row_end; j += row_step) rows.push_back(j);
    std::atomic<size_t> next{0};
    std::vector<Counters> cns((size_t)nthreads);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        Counters& cn = cns[(size_t)tid];
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= rows.size()) break;
            uint32_t j = rows[k];
            for (uint32_t i = 0; i < cam.c.image_width; ++i) {
                Col c = cast_ray(sc, cam, i, j, seed, cn);
                size_t o = ((size_t)j * cam.c.image_width + i) * 3;
                if (out_rgb) {
                    out_rgb[o] = c.r;
                    out_rgb[o + 1] = c.g;
                    out_rgb[o + 2] = c.b;
                }
                if (out_rgb8) {
                    uint32_t b[3];
                    col_bytes(c, b);
                    out_rgb8[o] = (uint8_t)b[0];
                    out_rgb8[o + 1] = (uint8_t)b[1];
                    out_rgb8[o + 2] = (uint8_t)b[2];
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (st) {
        memset(st, 0, sizeof(*st));
        for (auto& c : cns) {
            st->rays += c.rays;
            st->node_tests += c.node;
            st->sphere_tests += c.sph;
            st->tri_tests += c.tri;
            st->quad_tests += c.quad;
        }
        st->samples = (uint64_t)rows.size() * cam.c.image_width * cam.c.samples;
        st->seconds = std::chrono::duration<double>(t1 - t0).count();
        st->threads = nthreads;
    }
    return 0;
}

/* Ray batches for the id-parity tests.
 *  kind 0: pixel-centre primary rays (offset 0, no lens, time = frame time), n must be W*H
 *  kind 1: first-bounce rays: sample (pixel = k % (W*H), sample = k / (W*H)) of the camera stream,
 *          traced and scattered once; a ray that misses / is absorbed is replaced by its camera ray.
 *  kind 2: the camera sample rays themselves (jitter, lens, shutter).
 *  kind 3: as kind 1 with the pixels spread over the whole image (pixel = k * 2654435761 mod W*H), so that a batch
 *          smaller than the image still sees the whole scene. */
int orc_gen_rays(const OrcScene* h, const CrCamera* cam_in, uint64_t seed, int kind, size_t n, double* out) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    Cam cam;
    cam.c = *cam_in;
    const size_t wh = (size_t)cam.c.image_width * cam.c.image_height;
    Counters cn;
    for (size_t k = 0; k < n; ++k) {
        uint32_t pixel = (uint32_t)(k % wh), sample = (uint32_t)(k / wh);
        if (kind == 3) pixel = (uint32_t)(((uint64_t)k * 2654435761ull) % wh);
        uint32_t i = pixel % cam.c.image_width, j = pixel / cam.c.image_width;
        Ray r;
        if (kind == 0) {
            double t = cam.current_time();
            V3 cc = cam.get_from(t);
            V3 ps = cam.get_pixel_pos(i, j, V3{0.0, 0.0, 0.0}, t);
            r = {cc, sub(ps, cc), t};
        } else {
            Rng g(seed, pixel, sample, 0);
            r = cam.sample_ray(i, j, g);
            if (kind == 1 || kind == 3) {
                if (!sc.built) return CR_ERR_STATE;
                HitRecord hr;
                if (world_hit(sc, r, Interval{0.001, INF}, hr, cn)) {
                    Col att;
                    Ray s;
                    Rng g1(seed, pixel, sample, 1);
                    if (sc.mats[hr.mat].kind != CR_MAT_EMISSIVE && scatter(sc, r, hr, att, s, g1, !sc.has_emissive)) r = s;
                }
            }
        }
        double* p = out + 7 * k;
        p[0] = r.o.x; p[1] = r.o.y; p[2] = r.o.z;
        p[3] = r.d.x; p[4] = r.d.y; p[5] = r.d.z;
        p[6] = r.tm;
    }
    return 0;
}

/* ---- known-answer hooks for the reference's own unit tests (tests/test_kat_reference.py) ---- */
void orc_kat_vec(int op, const double* a, const double* b, double* out) {
    V3 x = {a[0], a[1], a[2]};
    V3 y = b ? V3{b[0], b[1], b[2]} : V3{0, 0, 0};
    V3 r = {0, 0, 0};
    switch (op) {
        case 0: r = neg(x); break;
        case 1: r = add(x, y); break;
        case 2: r = {dot(x, y), 0, 0}; break;
        case 3: r = cross(x, y); break;
        case 4: r = {len(x), 0, 0}; break;
        case 5: r = sub(x, y); break;
        case 6: r = unit(x); break;
        case 7: r = reflect(x, y); break;
        case 8: r = divs(x, y.x); break;
    }
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* Color::new range check (utils.rs:345-351): 1 = valid, 0 = would panic */
int orc_kat_color_valid(double r, double g, double b) {
    return (r <= 1.0 && g <= 1.0 && b <= 1.0 && r >= 0.0 && g >= 0.0 && b >= 0.0) ? 1 : 0;
}
void orc_kat_color_bytes(const double* c, uint32_t* out) { col_bytes(Col{c[0], c[1], c[2]}, out); }
void orc_kat_color_neg(const double* c, double* out) {
    Col r = col_neg(Col{c[0], c[1], c[2]});
    out[0] = r.r; out[1] = r.g; out[2] = r.b;
}
void orc_kat_color_add(const double* a, const double* b, double* out) {
    Col r = col_add(Col{a[0], a[1], a[2]}, Col{b[0], b[1], b[2]}, true);
    out[0] = r.r; out[1] = r.g; out[2] = r.b;
}
void orc_kat_color_scale(double s, const double* a, double* out) {
    Col r = col_scale(s, Col{a[0], a[1], a[2]}, true);
    out[0] = r.r; out[1] = r.g; out[2] = r.b;
}
/* average_samples, ray_casting.rs:154-173 */
void orc_kat_average(const double* cols, size_t n, double* out) {
    double r = 0, g = 0, b = 0;
    for (size_t i = 0; i < n; ++i) {
        r += cols[3 * i];
        g += cols[3 * i + 1];
        b += cols[3 * i + 2];
    }
    r /= (double)n;
    g /= (double)n;
    b /= (double)n;
    out[0] = r; out[1] = g; out[2] = b;
}
/* Interval ops: 0 size, 1 contains, 2 surrounds, 3 is_greater, 4 is_less, 5 proportion */
double orc_kat_interval(int op, double lo, double hi, double x) {
    Interval iv = {lo, hi};
    switch (op) {
        case 0: return iv.size();
        case 1: return iv.contains(x) ? 1.0 : 0.0;
        case 2: return iv.surrounds(x) ? 1.0 : 0.0;
        case 3: return iv.is_greater(x) ? 1.0 : 0.0;
        case 4: return iv.is_less(x) ? 1.0 : 0.0;
        default: return iv.proportion(x);
    }
}
/* Degrees/Radians, utils.rs:8-65 */
double orc_kat_deg_to_rad(double d) { return d * PI / 180.0; }
double orc_kat_rad_to_deg(double r) { return r * 180.0 / PI; }
void orc_kat_ray_at(const double* o, const double* d, double t, double* out) {
    Ray r = {{o[0], o[1], o[2]}, {d[0], d[1], d[2]}, 0.0};
    V3 p = r.at(t);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void orc_kat_point_at(const double* init, const CrKeyframe* keys, uint32_t n, double t, double* out) {
    V3 p = point_at(init, keys, n, t);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox(ctr, key, out); }
/* first n uniforms of stream (seed; pixel, sample, bounce) */
void orc_rng_stream(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, size_t n, double* out) {
    Rng g(seed, pixel, sample, bounce);
    for (size_t i = 0; i < n; ++i) out[i] = g.next();
}
/* single primitive hit tests (analytic KATs): kind 0 sphere [c,r], 1 triangle [a,b,c], 2 quad; ray [7] */
int orc_kat_prim_hit(int kind, const double* prim, const double* ray, double tmin, double tmax, CrHit* out) {
    Scene sc;
    OrcScene* h = reinterpret_cast<OrcScene*>(&sc);
    int32_t m = 0;
    if (kind == 0) orc_scene_add_spheres(h, prim, &m, nullptr, 1);
    else if (kind == 1) orc_scene_add_triangles(h, prim, &m, nullptr, 1);
    else orc_scene_add_quads(h, prim, &m, nullptr, 1);
    Ray r = {{ray[0], ray[1], ray[2]}, {ray[3], ray[4], ray[5]}, ray[6]};
    HitRecord hr;
    Counters cn;
    bool got = obj_hit(sc, sc.elements[0], r, Interval{tmin, tmax}, hr, cn);
    fill_hit(hr, got, *out);
    return got ? 1 : 0;
}
/* Aabb::hit KAT: box = [xmin,xmax,ymin,ymax,zmin,zmax] */
int orc_kat_aabb_hit(const double* box, const double* ray, double tmin, double tmax) {
    Aabb b = {{box[0], box[1]}, {box[2], box[3]}, {box[4], box[5]}};
    Ray r = {{ray[0], ray[1], ray[2]}, {ray[3], ray[4], ray[5]}, ray[6]};
    return aabb_hit(b, r, Interval{tmin, tmax}) ? 1 : 0;
}
/* scatter KAT: one scatter event from an explicit hit; returns 1 if scattered */
int orc_kat_scatter(const OrcScene* h, const double* ray_in, const CrHit* hit, uint64_t seed, uint32_t pixel,
                    uint32_t sample, uint32_t bounce, double* att, double* ray_out) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    Ray r = {{ray_in[0], ray_in[1], ray_in[2]}, {ray_in[3], ray_in[4], ray_in[5]}, ray_in[6]};
    HitRecord rec;
    rec.loc = {hit->p[0], hit->p[1], hit->p[2]};
    rec.normal = {hit->n[0], hit->n[1], hit->n[2]};
    rec.mat = hit->material;
    rec.t = hit->t;
    rec.u = hit->u;
    rec.v = hit->v;
    rec.front_face = hit->front_face != 0;
    Col a = {0, 0, 0};
    Ray s = r;
    Rng g(seed, pixel, sample, bounce);
    bool ok = scatter(sc, r, rec, a, s, g, !sc.has_emissive);
    att[0] = a.r; att[1] = a.g; att[2] = a.b;
    ray_out[0] = s.o.x; ray_out[1] = s.o.y; ray_out[2] = s.o.z;
    ray_out[3] = s.d.x; ray_out[4] = s.d.y; ray_out[5] = s.d.z;
    ray_out[6] = s.tm;
    return ok ? 1 : 0;
}
void orc_kat_tex_value(const OrcScene* h, int tex, double u, double v, const double* p, double* out) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    Col c = tex_value(sc, tex, u, v, V3{p[0], p[1], p[2]});
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void orc_kat_sky(const OrcScene* h, const double* dir, double* out) {
    const Scene& sc = *reinterpret_cast<const Scene*>(h);
    Ray r = {{0, 0, 0}, {dir[0], dir[1], dir[2]}, 0.0};
    Col c = sky_color(sc, r, !sc.has_emissive);
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

} /* extern "C" */
