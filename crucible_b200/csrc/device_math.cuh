// device_math.cuh — Vec3 / Color / Interval arithmetic in the reference's operation order
// (src/utils.rs), Philox4x32-10, and the per-primitive hit tests.  Templated on R = double
// (bit-faithful path; its translation unit is compiled with -fmad=false so no DFMA is formed)
// or float (fast path).
#pragma once
#include "common.cuh"

namespace crb {

template <typename R> struct Num;
template <> struct Num<double> {
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }  // IEEE rn
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double floor_(double x) { return floor(x); }
    static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double acos_(double x) { return acos(x); }
    static __device__ __forceinline__ double asin_(double x) { return asin(x); }
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double eps() { return 2.220446049250313e-16; }  // f64::EPSILON
    static __device__ __forceinline__ double pi() { return 3.14159265358979323846; }
};
template <> struct Num<float> {
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float floor_(float x) { return floorf(x); }
    static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float acos_(float x) { return acosf(x); }
    static __device__ __forceinline__ float asin_(float x) { return asinf(x); }
    static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float eps() { return 1.1920929e-7f; }
    static __device__ __forceinline__ float pi() { return 3.14159265358979323846f; }
};

// ---- utils.rs Point3/Vec3 (operation order is the contract) --------------------------------------
template <typename R> __device__ __forceinline__ V3<R> vneg(V3<R> a) { return {-a.x, -a.y, -a.z}; }                 // :248-257
template <typename R> __device__ __forceinline__ V3<R> vadd(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // :283-293
template <typename R> __device__ __forceinline__ V3<R> vsub(V3<R> a, V3<R> b) { return vadd(a, vneg(b)); }          // :295-301
template <typename R> __device__ __forceinline__ V3<R> vmul(R s, V3<R> v) { return {s * v.x, s * v.y, s * v.z}; }   // :303-311
template <typename R> __device__ __forceinline__ V3<R> vdiv(V3<R> v, R s) { return vmul(R(1) / s, v); }             // :325-331
template <typename R> __device__ __forceinline__ R vlen2(V3<R> v) { return v.x * v.x + v.y * v.y + v.z * v.z; }     // :183-186
template <typename R> __device__ __forceinline__ R vdot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // :194-199
template <typename R> __device__ __forceinline__ V3<R> vcross(V3<R> v, V3<R> o) {                                    // :201-212
    return {v.y * o.z - v.z * o.y, v.z * o.x - v.x * o.z, v.x * o.y - v.y * o.x};
}
template <typename R> __device__ __forceinline__ V3<R> vunit(V3<R> v) { return vdiv(v, Num<R>::sqrt_(vlen2(v))); }  // :215-218
template <typename R> __device__ __forceinline__ bool vnear_zero(V3<R> v) {                                         // :189-192
    const R tol = R(1e-8);
    return Num<R>::abs_(v.x) < tol && Num<R>::abs_(v.y) < tol && Num<R>::abs_(v.z) < tol;
}
template <typename R> __device__ __forceinline__ V3<R> vreflect(V3<R> v, V3<R> n) {                                  // :151-153
    return vsub(v, vmul(R(2) * vdot(v, n), n));
}
template <typename R> __device__ __forceinline__ V3<R> vrefract(V3<R> v, V3<R> n, R eta) {                           // :159-165
    R cos_theta = Num<R>::min_(vdot(vneg(v), n), R(1));
    V3<R> perp = vmul(eta, vadd(v, vmul(cos_theta, n)));
    V3<R> par = vmul(-(Num<R>::sqrt_(Num<R>::abs_(R(1) - vlen2(perp)))), n);
    return vadd(perp, par);
}
template <typename R> __device__ __forceinline__ R rclamp(R x, R lo, R hi) {  // Rust f64::clamp
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

// ---- utils.rs Color: every operator clamps to [0,1] (utils.rs:487-603) ---------------------------
template <typename R> __device__ __forceinline__ R c01(R x, bool cl) { return cl ? rclamp(x, R(0), R(1)) : x; }
template <typename R> __device__ __forceinline__ V3<R> col_neg(V3<R> c) {  // :445-459 (hilo complement)
    R mn = (c.x < c.y ? c.x : c.y);
    mn = (mn < c.z ? mn : c.z);
    R mx = (c.x > c.y ? c.x : c.y);
    mx = (mx > c.z ? mx : c.z);
    R k = mn + mx;
    return {Num<R>::abs_(k - c.x), Num<R>::abs_(k - c.y), Num<R>::abs_(k - c.z)};
}
template <typename R> __device__ __forceinline__ V3<R> col_add(V3<R> a, V3<R> b, bool cl) {  // :516-529
    return {c01(a.x + b.x, cl), c01(a.y + b.y, cl), c01(a.z + b.z, cl)};
}
template <typename R> __device__ __forceinline__ V3<R> col_scale(R s, V3<R> c, bool cl) {  // :558-574
    V3<R> m = (s < R(0)) ? col_neg(c) : c;
    R p = Num<R>::abs_(s);
    return {c01(p * m.x, cl), c01(p * m.y, cl), c01(p * m.z, cl)};
}
template <typename R> __device__ __forceinline__ V3<R> col_mul(V3<R> a, V3<R> b, bool cl) {  // :576-590
    return {c01(a.x * b.x, cl), c01(a.y * b.y, cl), c01(a.z * b.z, cl)};
}
template <typename R> __device__ __forceinline__ V3<R> col_div(V3<R> c, R rhs, bool cl) {  // :592-601
    V3<R> inv = (rhs < R(0)) ? col_neg(c) : c;
    return col_scale(R(1) / Num<R>::abs_(rhs), inv, cl);
}

// ---- Philox4x32-10, keyed (seed; pixel, sample, bounce, block) -----------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                        uint32_t k1, uint32_t out[4]) {
#pragma unroll
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

template <typename R> struct Rng;
// f64: two 53-bit uniforms per block, identical to the oracle's stream (path-by-path comparable)
template <> struct Rng<double> {
    uint32_t k0, k1, c0, c1, c2, c3;
    uint32_t b2, b3;
    int have;
    __device__ __forceinline__ Rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0(pixel), c1(sample), c2(bounce), c3(0), b2(0), b3(0), have(0) {}
    __device__ __forceinline__ double next() {
        uint64_t x;
        if (have == 0) {
            uint32_t o[4];
            philox4x32_10(c0, c1, c2, c3, k0, k1, o);
            c3++;
            b2 = o[2];
            b3 = o[3];
            have = 1;
            x = ((uint64_t)o[1] << 32) | o[0];
        } else {
            have = 0;
            x = ((uint64_t)b3 << 32) | b2;
        }
        return (double)(x >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double range(double lo, double hi) { return lo + (hi - lo) * next(); }
};
// f32: four 24-bit uniforms per block (a different, equally valid stream)
template <> struct Rng<float> {
    uint32_t k0, k1, c0, c1, c2, c3;
    uint32_t b[4];
    int have;
    __device__ __forceinline__ Rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0(pixel), c1(sample), c2(bounce), c3(0), have(0) {}
    __device__ __forceinline__ float next() {
        if (have == 0) {
            philox4x32_10(c0, c1, c2, c3, k0, k1, b);
            c3++;
            have = 4;
        }
        uint32_t w = (have == 4) ? b[0] : (have == 3) ? b[1] : (have == 2) ? b[2] : b[3];
        have--;
        return (float)(w >> 8) * (1.0f / 16777216.0f);
    }
    __device__ __forceinline__ float range(float lo, float hi) { return lo + (hi - lo) * next(); }
};

template <typename R> __device__ __forceinline__ V3<R> random_unit_vector(Rng<R>& g) {  // utils.rs:128-138
    for (;;) {
        R x = g.range(R(-1), R(1));
        R y = g.range(R(-1), R(1));
        R z = g.range(R(-1), R(1));
        V3<R> p = {x, y, z};
        R l2 = vlen2(p);
        if (R(1e-160) < l2 && l2 <= R(1)) return vdiv(p, Num<R>::sqrt_(l2));
    }
}
template <> __device__ __forceinline__ V3<float> random_unit_vector<float>(Rng<float>& g) {
    for (;;) {
        float x = g.range(-1.f, 1.f), y = g.range(-1.f, 1.f), z = g.range(-1.f, 1.f);
        V3<float> p = {x, y, z};
        float l2 = vlen2(p);
        if (1e-30f < l2 && l2 <= 1.f) return vdiv(p, sqrtf(l2));
    }
}
template <typename R> __device__ __forceinline__ V3<R> random_in_unit_disk(Rng<R>& g) {  // utils.rs:112-126
    for (;;) {
        R x = g.range(R(-1), R(1));
        R y = g.range(R(-1), R(1));
        V3<R> p = {x, y, R(0)};
        if (vlen2(p) < R(1)) return p;
    }
}

// ---- Aabb::hit, src/objects/bvh.rs:96-132, comparison form (defines NaN behaviour) ---------------
// inv = 1.0 / d per axis is a pure function of the ray, so it is hoisted out of the node loop.
template <typename R>
__device__ __forceinline__ bool slab_axis(R bmin, R bmax, R o, R inv, R& tmin, R& tmax) {
    R t0 = (bmin - o) * inv;
    R t1 = (bmax - o) * inv;
    R nmin, nmax;
    if (t0 < t1) {
        nmin = (t0 > tmin) ? t0 : tmin;
        nmax = (t1 < tmax) ? t1 : tmax;
    } else {
        nmin = (t1 > tmin) ? t1 : tmin;
        nmax = (t0 < tmax) ? t0 : tmax;
    }
    tmin = nmin;
    tmax = nmax;
    return !(tmax <= tmin);
}
template <typename R>
__device__ __forceinline__ bool aabb_hit(const NodeRec<R>& n, V3<R> o, V3<R> inv, R tmin, R tmax) {
    if (!slab_axis(n.xmin, n.xmax, o.x, inv.x, tmin, tmax)) return false;
    if (!slab_axis(n.ymin, n.ymax, o.y, inv.y, tmin, tmax)) return false;
    return slab_axis(n.zmin, n.zmax, o.z, inv.z, tmin, tmax);
}

// ---- Sphere::hit, src/objects/sphere.rs:61-105: returns the accepted root ------------------------
template <typename R>
__device__ __forceinline__ bool sphere_hit_t(const SphereRec<R>& s, V3<R> o, V3<R> d, R a, R tmin, R tmax, R& t) {
    V3<R> oc = vsub(V3<R>{s.cx, s.cy, s.cz}, o);
    R h = vdot(d, oc);
    R c = vlen2(oc) - s.r * s.r;
    R disc = h * h - a * c;
    if (disc < R(0)) return false;
    R sq = Num<R>::sqrt_(disc);
    R root = (h - sq) / a;
    if (!(tmin < root && root < tmax)) {
        root = (h + sq) / a;
        if (!(tmin < root && root < tmax)) return false;
    }
    t = root;
    return true;
}
// ---- Triangle::hit, src/objects/triangle.rs:86-140 (e1, e2 precomputed with the same ops) --------
template <typename R>
__device__ __forceinline__ bool tri_hit_t(const TriRec<R>& tr, V3<R> o, V3<R> d, R tmin, R tmax, R& t) {
    V3<R> e1 = {tr.e1x, tr.e1y, tr.e1z}, e2 = {tr.e2x, tr.e2y, tr.e2z};
    V3<R> pv = vcross(d, e2);
    R det = vdot(e1, pv);
    if (det > -Num<R>::eps() && det < Num<R>::eps()) return false;
    R inv = R(1) / det;
    V3<R> s = vsub(o, V3<R>{tr.ax, tr.ay, tr.az});
    R u = inv * vdot(s, pv);
    if (!(R(0) <= u && u <= R(1))) return false;
    V3<R> q = vcross(s, e1);
    R v = inv * vdot(d, q);
    if (v < R(0) || u + v > R(1)) return false;
    R tt = inv * vdot(e2, q);
    if (!(tmin < tt && tt < tmax)) return false;
    t = tt;
    return true;
}
// ---- EXTENSION quad (same definition as oracle/oracle.cpp quad_hit) -------------------------------
template <typename R>
__device__ __forceinline__ bool quad_hit_t(const QuadRec<R>& q, V3<R> o, V3<R> d, R tmin, R tmax, R& t, R& alpha, R& beta) {
    V3<R> n = {q.nx, q.ny, q.nz};
    R denom = vdot(n, d);
    if (Num<R>::abs_(denom) < R(1e-8)) return false;
    R tt = (q.d - vdot(n, o)) / denom;
    if (!(tmin < tt && tt < tmax)) return false;
    V3<R> p = vadd(o, vmul(tt, d));
    V3<R> ph = vsub(p, V3<R>{q.qx, q.qy, q.qz});
    V3<R> w = {q.wx, q.wy, q.wz};
    alpha = vdot(w, vcross(ph, V3<R>{q.vx, q.vy, q.vz}));
    beta = vdot(w, vcross(V3<R>{q.ux, q.uy, q.uz}, ph));
    if (!(R(0) <= alpha && alpha <= R(1)) || !(R(0) <= beta && beta <= R(1))) return false;
    t = tt;
    return true;
}

// Sphere / triangle records at ray time tm: the stored record for static primitives (and static scenes),
// the timelines evaluated at tm otherwise (Sphere::hit sphere.rs:67-70, Triangle::hit triangle.rs:91-100).
// ANIM = false is the build static scenes run: no track lookups, no extra registers in the hot kernels.
template <typename R, bool ANIM>
__device__ __forceinline__ SphereRec<R> sphere_at(const DevScene<R>& sc, uint32_t idx, R tm) {
    SphereRec<R> s = ldg_rec<sizeof(SphereRec<R>) / 16>(sc.spheres + idx);
    if constexpr (!ANIM) return s;
    if (sc.sphere_track != nullptr) {
        const AnimTrack tr = sc.sphere_track[idx];
        if (tr.count) {
            R p[3] = {s.cx, s.cy, s.cz};
            anim_eval<R>(sc.anim_keys, tr.first, tr.count, tm, p, s.r);
            s.cx = p[0]; s.cy = p[1]; s.cz = p[2];
        }
    }
    return s;
}
template <typename R, bool ANIM>
__device__ __forceinline__ TriRec<R> tri_at(const DevScene<R>& sc, uint32_t idx, R tm) {
    TriRec<R> t = ldg_rec<sizeof(TriRec<R>) / 16>(sc.tris + idx);
    if constexpr (!ANIM) return t;
    if (sc.tri_track != nullptr) {
        const AnimTrack ta = sc.tri_track[3 * idx], tb = sc.tri_track[3 * idx + 1], tc = sc.tri_track[3 * idx + 2];
        if (ta.count | tb.count | tc.count) {
            // the record holds a, e1 = b - a, e2 = c - a of the construction vertices; b and c themselves are in the
            // animated-vertex table (exact copies: b cannot be recovered from a + e1 bit for bit)
            const double* v = sc.tri_anim_verts + 9ull * sc.tri_anim_slot[idx];
            R a[3] = {(R)v[0], (R)v[1], (R)v[2]}, b[3] = {(R)v[3], (R)v[4], (R)v[5]}, c[3] = {(R)v[6], (R)v[7], (R)v[8]};
            R unused = R(1);
            anim_eval<R>(sc.anim_keys, ta.first, ta.count, tm, a, unused);
            anim_eval<R>(sc.anim_keys, tb.first, tb.count, tm, b, unused);
            anim_eval<R>(sc.anim_keys, tc.first, tc.count, tm, c, unused);
            t.ax = a[0]; t.ay = a[1]; t.az = a[2];
            t.e1x = b[0] + (-a[0]); t.e1y = b[1] + (-a[1]); t.e1z = b[2] + (-a[2]);  // triangle.rs:99-100
            t.e2x = c[0] + (-a[0]); t.e2y = c[1] + (-a[1]); t.e2z = c[2] + (-a[2]);
        }
    }
    return t;
}
template <typename R, bool ANIM>
__device__ __forceinline__ bool sphere_is_animated(const DevScene<R>& sc, uint32_t idx) {
    if constexpr (!ANIM) return false;
    return sc.sphere_track != nullptr && sc.sphere_track[idx].count != 0u;
}

// ---- HitRecord (src/objects/mod.rs:21-87) rebuilt from (ref, t) -----------------------------------
template <typename R> struct HitInfo {
    V3<R> p, n;
    R u, v;
    bool front;
    int32_t material, mat_kind, prim_index, obj_id;
};
// HitRecord geometry rebuilt from (ref, t).  want_uv = false skips get_sphere_uv (acos + atan2): the
// texture coordinates only reach image textures (solid_color.rs:25-27 and checker_texture.rs:39-51
// ignore u, v), so skipping them cannot change a result.
template <typename R, bool ANIM>
__device__ __forceinline__ HitInfo<R> finalize_geom(const DevScene<R>& sc, uint32_t ref, R t, V3<R> o, V3<R> d, bool want_uv, R tm) {
    HitInfo<R> h;
    const uint32_t kind = ref_kind(ref), idx = ref_index(ref);
    h.material = h.mat_kind = h.prim_index = h.obj_id = -1;
    h.p = vadd(o, vmul(t, d));  // Ray::at, ray_casting.rs:53-59
    V3<R> n;
    if (kind == CR_PRIM_SPHERE) {
        SphereRec<R> s = sphere_at<R, ANIM>(sc, idx, tm);
        n = vdiv(vsub(h.p, V3<R>{s.cx, s.cy, s.cz}), s.r);  // sphere.rs:96
        h.u = h.v = R(0);
        if (want_uv) {
            R theta = Num<R>::acos_(-n.y);  // sphere.rs:41-46
            R phi = Num<R>::atan2_(-n.z, n.x) + Num<R>::pi();
            h.u = phi / (R(2) * Num<R>::pi());
            h.v = theta / Num<R>::pi();
        }
    } else if (kind == CR_PRIM_TRIANGLE) {
        TriRec<R> tr = tri_at<R, ANIM>(sc, idx, tm);
        n = vunit(vcross(V3<R>{tr.e1x, tr.e1y, tr.e1z}, V3<R>{tr.e2x, tr.e2y, tr.e2z}));  // triangle.rs:124 + safe_new
        h.u = R(0);  // triangle.rs:133-134
        h.v = R(0);
    } else {
        QuadRec<R> q = ldg_rec<sizeof(QuadRec<R>) / 16>(sc.quads + idx);
        n = {q.nx, q.ny, q.nz};
        V3<R> ph = vsub(h.p, V3<R>{q.qx, q.qy, q.qz});
        V3<R> w = {q.wx, q.wy, q.wz};
        h.u = vdot(w, vcross(ph, V3<R>{q.vx, q.vy, q.vz}));
        h.v = vdot(w, vcross(V3<R>{q.ux, q.uy, q.uz}, ph));
    }
    h.front = vdot(d, n) < R(0);  // objects/mod.rs:47-48
    h.n = h.front ? n : vneg(n);
    return h;
}
// full record incl. the ids of SURVEY 8b (cr_trace_batch)
template <typename R, bool ANIM>
__device__ __forceinline__ HitInfo<R> finalize_hit(const DevScene<R>& sc, uint32_t ref, R t, V3<R> o, V3<R> d, R tm) {
    HitInfo<R> h = finalize_geom<R, ANIM>(sc, ref, t, o, d, true, tm);
    const uint32_t kind = ref_kind(ref);
    const PrimMeta* mp = kind == CR_PRIM_SPHERE ? sc.meta[0] : (kind == CR_PRIM_TRIANGLE ? sc.meta[1] : sc.meta[2]);
    const PrimMeta m = mp[ref_index(ref)];
    h.material = m.material;
    h.mat_kind = m.mat_kind & MATKIND_MASK;
    h.prim_index = m.prim_index;
    h.obj_id = m.obj_id;
    return h;
}

// ---- closest hit: BVHWrapper::hit, src/objects/bvhwrapper.rs:97-126 -------------------------------------
// The recursion "left with (tmin,tmax), right with (tmin, left.t or tmax), prefer right" is a DFS with
// ONE running closest-t: a later primitive only wins when strictly closer, inner boxes are tested with
// the running interval at the time they are entered, leaves are tested with no box of their own.
//
// Branch-free box test for REGULAR rays (every origin/direction component finite, every 1/d finite and
// non-zero).  For such rays t0, t1 are never NaN, so the reference's comparison form reduces to
//   lo = max(tmin, min(t0,t1) per axis), hi = min(tmax, max(t0,t1) per axis), hit <=> hi > lo
// and `t0 < t1` is decided by the sign of 1/d (equal values make both arms identical).  The per-axis
// early exits of bvh.rs:127-129 are implied: lo only grows and hi only shrinks, so an intermediate
// `hi <= lo` stays true to the end.  Decisions are therefore IDENTICAL to aabb_hit(); irregular rays
// (axis parallel, NaN, inf) take aabb_hit() itself.
template <typename R>
__device__ __forceinline__ bool aabb_hit_regular(const NodeRec<R>& n, V3<R> o, V3<R> inv, bool px, bool py, bool pz, R tmin,
                                                 R tmax) {
    const R x0 = (n.xmin - o.x) * inv.x, x1 = (n.xmax - o.x) * inv.x;
    const R y0 = (n.ymin - o.y) * inv.y, y1 = (n.ymax - o.y) * inv.y;
    const R z0 = (n.zmin - o.z) * inv.z, z1 = (n.zmax - o.z) * inv.z;
    const R lx = px ? x0 : x1, hx = px ? x1 : x0;
    const R ly = py ? y0 : y1, hy = py ? y1 : y0;
    const R lz = pz ? z0 : z1, hz = pz ? z1 : z0;
    R lo = (lx > tmin) ? lx : tmin;
    R hi = (hx < tmax) ? hx : tmax;
    lo = (ly > lo) ? ly : lo;
    hi = (hy < hi) ? hy : hi;
    lo = (lz > lo) ? lz : lo;
    hi = (hz < hi) ? hz : hi;
    return hi > lo;
}
template <typename R>
__device__ __forceinline__ bool is_finite(R x) {
    return Num<R>::abs_(x) < Num<R>::inf();
}

// ---- conservative f32 filters (f64 path) ----------------------------------------------------------
// BOX.  The reference's decision at a node is  min(best_t, hi_x, hi_y, hi_z) > max(tmin, lo_x, lo_y, lo_z)
// evaluated in f64.  The same expression evaluated in f32 on the outward-rounded f32 copy of the box
// differs from the f64 values by at most
//     |t32 - t64| <= |inv| (|b| 2^-23 + |o| 2^-24)(1 + 2^-22) + 3 * 2^-24 |t64|
// (rounding of b, o and inv to f32, one subtraction, one multiplication).  With B an upper bound of the
// box's |coordinates|, e = max over axes of |inv|(B 2^-23 + |o| 2^-24) is a per-ray constant and
//     E = 2.5 e + 2^-21 (|lo32| + |hi32|)
// bounds the error of (hi - lo).  So  hi32 - lo32 > E  =>  the f64 test passes,  hi32 - lo32 < -E  =>  it
// fails, and only the sliver in between is re-decided by the exact f64 test.  Two bounds are kept per ray:
// e_big for B = the root box (book1: 2000, because of the r=1000 ground sphere) and e_small for the
// B_small that covers ~90 % of the nodes; bit 27 of a node's first word says which one applies.
//
// SPHERE.  disc = h^2 - a (|oc|^2 - r^2) evaluated in f32 from f32 copies of c, r, o, d differs from the
// f64 value by at most  2^-17.5 a (|c|^2 + |o|^2 + r^2)  (derivation in DESIGN.md section 5.1), so
// disc32 < -1.0e-5 a32 (|c|^2+|o|^2+r^2)  =>  the f64 discriminant is negative  =>  Sphere::hit returns None
// (sphere.rs:79-81).  A leaf node whose primitives are all definite misses is not visited at all.
//
// In both cases the decision taken is ALWAYS the reference's decision; the filters only save work.
struct FilterRay {
    float ox, oy, oz, ix, iy, iz, dx, dy, dz;
    float e_small, e_big, o2;
    // derived at lane refill (f64 path): the box filter evaluates every plane as ONE fused multiply-add,
    //   t = fma(b, inv, -(o * inv)),  near/far chosen by the sign of inv through a (inv, 0) / (0, inv) pair:
    //   near = fma(bmin, a0, fma(bmax, a1, -oi)),  far = fma(bmax, a0, fma(bmin, a1, -oi))
    // (fma(x, 0, c) == c exactly for finite x), which moves the six per-axis min/max out of the ALU pipe.
    float oix, oiy, oiz, ax0, ax1, ay0, ay1, az0, az1;
    bool ok;  // the ray is regular and representable in f32: filters may be used
    __device__ __forceinline__ void derive() {
        oix = ox * ix; oiy = oy * iy; oiz = oz * iz;
        ax0 = fmaxf(ix, 0.f); ax1 = fminf(ix, 0.f);
        ay0 = fmaxf(iy, 0.f); ay1 = fminf(iy, 0.f);
        az0 = fmaxf(iz, 0.f); az1 = fminf(iz, 0.f);
    }
};
static constexpr uint32_t BIGBOX_BIT = 1u << 27;
static constexpr uint32_t INDEX_MASK = 0x07FFFFFFu;

template <typename R>
__device__ __forceinline__ FilterRay make_filter_ray(V3<R> o, V3<R> d, V3<R> inv, R tmin, R tmax, float bsmall, float bmax) {
    FilterRay f;
    f.ox = (float)o.x; f.oy = (float)o.y; f.oz = (float)o.z;
    f.ix = (float)inv.x; f.iy = (float)inv.y; f.iz = (float)inv.z;
    f.dx = (float)d.x; f.dy = (float)d.y; f.dz = (float)d.z;
    const float ax = fabsf(f.ix), ay = fabsf(f.iy), az = fabsf(f.iz);
    const float u23 = 1.1920929e-7f;  // 2^-23
    const float oo_x = fabsf(f.ox) * u23, oo_y = fabsf(f.oy) * u23, oo_z = fabsf(f.oz) * u23;
    const float ks = bsmall * u23, kb = bmax * u23;  // B * 2^-23
    f.e_small = 2.5f * fmaxf(ax * (ks + oo_x), fmaxf(ay * (ks + oo_y), az * (ks + oo_z)));
    f.e_big = 2.5f * fmaxf(ax * (kb + oo_x), fmaxf(ay * (kb + oo_y), az * (kb + oo_z)));
    f.o2 = f.ox * f.ox + f.oy * f.oy + f.oz * f.oz;
    const bool regular = is_finite(o.x) && is_finite(o.y) && is_finite(o.z) && is_finite(inv.x) && is_finite(inv.y) &&
                         is_finite(inv.z) && inv.x != R(0) && inv.y != R(0) && inv.z != R(0) && !(tmin != tmin) && !(tmax != tmax);
    // skipped for irregular rays and when f32 cannot represent the ray (overflow / underflow)
    f.derive();
    f.ok = regular && f.e_big < 3.0e37f && ax > 1.0e-30f && ay > 1.0e-30f && az > 1.0e-30f && f.o2 < 1.0e30f &&
           fabsf(f.oix) < 3.0e37f && fabsf(f.oiy) < 3.0e37f && fabsf(f.oiz) < 3.0e37f;
    return f;
}
// 48 B image of a FilterRay, written next to every path record by the kernel that PRODUCES the ray (raygen /
// shade, where all lanes do it together) so that the trace kernel's lane refill is three 16 B loads instead of
// six f64 loads, three f64 reciprocals and the bound arithmetic executed by a handful of lanes.
struct __align__(16) FilterRec {
    float f[12];
};
__device__ __forceinline__ FilterRec pack_filter(const FilterRay& r) {
    FilterRec o;
    o.f[0] = r.ox; o.f[1] = r.oy; o.f[2] = r.oz; o.f[3] = r.ix; o.f[4] = r.iy; o.f[5] = r.iz;
    o.f[6] = r.dx; o.f[7] = r.dy; o.f[8] = r.dz; o.f[9] = r.e_small; o.f[10] = r.e_big;
    o.f[11] = r.ok ? r.o2 : -1.0f;  // |o|^2 >= 0: a negative value marks "filters off"
    return o;
}
__device__ __forceinline__ FilterRay unpack_filter(const FilterRec& o) {
    FilterRay r;
    r.ox = o.f[0]; r.oy = o.f[1]; r.oz = o.f[2]; r.ix = o.f[3]; r.iy = o.f[4]; r.iz = o.f[5];
    r.dx = o.f[6]; r.dy = o.f[7]; r.dz = o.f[8]; r.e_small = o.f[9]; r.e_big = o.f[10]; r.o2 = o.f[11];
    r.ok = o.f[11] >= 0.0f;
    r.derive();
    return r;
}
// f32 evaluation of the slab intervals: the f32 path's own box test
template <typename F>
__device__ __forceinline__ void slab32(const NodeRec<float>& n, const F& f, float tmin, float best, float& lo, float& hi) {
    const float x0 = (n.xmin - f.ox) * f.ix, x1 = (n.xmax - f.ox) * f.ix;
    const float y0 = (n.ymin - f.oy) * f.iy, y1 = (n.ymax - f.oy) * f.iy;
    const float z0 = (n.zmin - f.oz) * f.iz, z1 = (n.zmax - f.oz) * f.iz;
    lo = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    hi = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), best));
}
// Conservative filter of the f64 path: +1 = passes, -1 = fails, 0 = undecided (exact f64 test required).
// Every plane distance is one FFMA, t32 = fl(b32 * inv32 - oi32) with oi32 = fl(o32 * inv32).  With u = 2^-24:
// |b32 - b| <= 2u|b| (outward rounding), |inv32 - inv| <= u|inv|, |o32 - o| <= u|o|, two more roundings, so
//     |t32 - t64| <= 1.01 |inv32| (|b| + |o32|) 2^-23 + 2.01 u |t64|.
// e = 2.5 max_axis |inv32| (B + |o32|) 2^-23 (per ray, B = bound of the node class) and 2^-21 (|lo32| + |hi32|)
// therefore bound the error of hi - lo with margin to spare.
template <typename F>
__device__ __forceinline__ int filter_box(const NodeRec<float>& n, const F& f, float tmin, float best) {
    const float nx = __fmaf_rn(n.xmin, f.ax0, __fmaf_rn(n.xmax, f.ax1, -f.oix));
    const float fx = __fmaf_rn(n.xmax, f.ax0, __fmaf_rn(n.xmin, f.ax1, -f.oix));
    const float ny = __fmaf_rn(n.ymin, f.ay0, __fmaf_rn(n.ymax, f.ay1, -f.oiy));
    const float fy = __fmaf_rn(n.ymax, f.ay0, __fmaf_rn(n.ymin, f.ay1, -f.oiy));
    const float nz = __fmaf_rn(n.zmin, f.az0, __fmaf_rn(n.zmax, f.az1, -f.oiz));
    const float fz = __fmaf_rn(n.zmax, f.az0, __fmaf_rn(n.zmin, f.az1, -f.oiz));
    const float lo = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float hi = fminf(fminf(fx, fy), fminf(fz, best));
    const float diff = hi - lo;
    const float E = __fmaf_rn(fabsf(lo) + fabsf(hi), 4.7683716e-7f, (n.left & BIGBOX_BIT) ? f.e_big : f.e_small);
    if (diff > E) return 1;
    if (diff < -E) return -1;
    return 0;  // also NaN / inf
}
// Is the box, in ray parameter, wider on every axis than k times the filter's error band?  Then the band only matters
// where the ray grazes the box's boundary and the ambiguity does not repeat in all of its descendants (rare path of
// step_node: recomputes the planes instead of keeping them live in the node loop).
template <typename F>
__device__ __forceinline__ bool box_wider_than_band(const NodeRec<float>& n, const F& f, float tmin, float best, float k) {
    const float nx = __fmaf_rn(n.xmin, f.ax0, __fmaf_rn(n.xmax, f.ax1, -f.oix));
    const float fx = __fmaf_rn(n.xmax, f.ax0, __fmaf_rn(n.xmin, f.ax1, -f.oix));
    const float ny = __fmaf_rn(n.ymin, f.ay0, __fmaf_rn(n.ymax, f.ay1, -f.oiy));
    const float fy = __fmaf_rn(n.ymax, f.ay0, __fmaf_rn(n.ymin, f.ay1, -f.oiy));
    const float nz = __fmaf_rn(n.zmin, f.az0, __fmaf_rn(n.zmax, f.az1, -f.oiz));
    const float fz = __fmaf_rn(n.zmax, f.az0, __fmaf_rn(n.zmin, f.az1, -f.oiz));
    const float lo = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float hi = fminf(fminf(fx, fy), fminf(fz, best));
    const float E = __fmaf_rn(fabsf(lo) + fabsf(hi), 4.7683716e-7f, (n.left & BIGBOX_BIT) ? f.e_big : f.e_small);
    const float w = fminf(fminf(fx - nx, fy - ny), fz - nz);
    return E * k < w;  // false for NaN
}
// The part of a FilterRay the sphere pre-filter needs (kept out of the node loop's registers).
struct PreRay {
    float ox, oy, oz, dx, dy, dz, o2;
};
// true = Sphere::hit certainly returns None for this ray (f64 discriminant certainly negative)
__device__ __forceinline__ bool sphere_definite_miss(const SphereRec<float>& s, const PreRay& f) {
    const float ocx = s.cx - f.ox, ocy = s.cy - f.oy, ocz = s.cz - f.oz;
    const float a = f.dx * f.dx + f.dy * f.dy + f.dz * f.dz;
    const float h = f.dx * ocx + f.dy * ocy + f.dz * ocz;
    const float c2 = s.cx * s.cx + s.cy * s.cy + s.cz * s.cz;
    const float r2 = s.r * s.r;
    const float cq = (ocx * ocx + ocy * ocy + ocz * ocz) - r2;
    const float disc = h * h - a * cq;
    const float err = 1.0e-5f * a * (c2 + f.o2 + r2);
    return disc < -err;  // false for NaN / inf
}

// ---- warp-persistent closest-hit engine -------------------------------------------------------------
// Lane states
enum : int { ST_IDLE = 0, ST_NODE = 1, ST_EXACT = 2, ST_LEAF = 3, ST_DONE = 4 };

// What the node loop keeps in registers per lane: the constants of the box test and nothing else.
template <typename R> struct NodeRay;
template <> struct NodeRay<double> {  // conservative filter: one FFMA per plane (filter_box)
    float oix, oiy, oiz, ax0, ax1, ay0, ay1, az0, az1, e_small, e_big;
    __device__ __forceinline__ void set(const FilterRay& f) {
        oix = f.oix; oiy = f.oiy; oiz = f.oiz; ax0 = f.ax0; ax1 = f.ax1; ay0 = f.ay0; ay1 = f.ay1; az0 = f.az0; az1 = f.az1;
        e_small = f.e_small; e_big = f.e_big;
    }
};
template <> struct NodeRay<float> {  // the f32 path's own box test (slab32)
    float ox, oy, oz, ix, iy, iz;
    __device__ __forceinline__ void set(const FilterRay& f) { ox = f.ox; oy = f.oy; oz = f.oz; ix = f.ix; iy = f.iy; iz = f.iz; }
};

// Everything else a lane carries (closest hit, the ray's index, the f32 ray of the sphere pre-filter, the
// R-precision ray of the exact / leaf steps) sits behind a Store.  RegStore = registers (k_tail: one thread
// follows one path).  SmemStore = a per-CTA shared-memory table, SoA over the CTA's lanes so every access
// is conflict free: the trace kernel is latency bound (ncu: 1.6 eligible warps per scheduler, long-scoreboard
// the top stall), so the registers this frees buy resident warps, and the exact / leaf steps read their ray
// at shared-memory latency instead of re-reading the path record from L2 / HBM.
template <typename R>
struct RegStore {
    R bt;
    uint32_t bref, idx;
    PreRay pre;
    V3<R> ro, rd;
    R tm;
    __device__ __forceinline__ R time() const { return tm; }
    __device__ __forceinline__ void set_time(R t) { tm = t; }
    __device__ __forceinline__ R best_t() const { return bt; }
    __device__ __forceinline__ uint32_t best_ref() const { return bref; }
    __device__ __forceinline__ void set_best(R t, uint32_t ref) { bt = t; bref = ref; }
    __device__ __forceinline__ uint32_t my() const { return idx; }
    __device__ __forceinline__ void set_my(uint32_t k) { idx = k; }
    __device__ __forceinline__ void set_pre(const FilterRay& f) { pre = {f.ox, f.oy, f.oz, f.dx, f.dy, f.dz, f.o2}; }
    __device__ __forceinline__ PreRay get_pre() const { return pre; }
    __device__ __forceinline__ void set_ray(V3<R> o, V3<R> d) { ro = o; rd = d; }
    __device__ __forceinline__ void get_ray(V3<R>& o, V3<R>& d) const { o = ro; d = rd; }
};
// (10 CTAs x (11.5 KB + 1 KB reserved) = 125 KB sits just inside the 132 KB shared-memory carve-out; the ray time
// of the animated builds lives in its own table so the static builds keep that L1 / shared split.)
template <typename R, int BLOCK, bool ANIM>
struct LaneSlots {
    R best_t[BLOCK];
    R tm[ANIM ? BLOCK : 1];  // ray time (positions animated primitives)
    R ray[6][BLOCK];
    float pre[7][BLOCK];
    uint32_t best_ref[BLOCK], my[BLOCK];
};
template <typename R, int BLOCK, bool ANIM>
struct SmemStore {
    LaneSlots<R, BLOCK, ANIM>* s;
    __device__ __forceinline__ R time() const { return s->tm[ANIM ? threadIdx.x : 0]; }
    __device__ __forceinline__ void set_time(R t) { s->tm[ANIM ? threadIdx.x : 0] = t; }
    __device__ __forceinline__ R best_t() const { return s->best_t[threadIdx.x]; }
    __device__ __forceinline__ uint32_t best_ref() const { return s->best_ref[threadIdx.x]; }
    __device__ __forceinline__ void set_best(R t, uint32_t ref) { s->best_t[threadIdx.x] = t; s->best_ref[threadIdx.x] = ref; }
    __device__ __forceinline__ uint32_t my() const { return s->my[threadIdx.x]; }
    __device__ __forceinline__ void set_my(uint32_t k) { s->my[threadIdx.x] = k; }
    __device__ __forceinline__ void set_pre(const FilterRay& f) {
        const int t = threadIdx.x;
        s->pre[0][t] = f.ox; s->pre[1][t] = f.oy; s->pre[2][t] = f.oz; s->pre[3][t] = f.dx; s->pre[4][t] = f.dy; s->pre[5][t] = f.dz;
        s->pre[6][t] = f.o2;
    }
    __device__ __forceinline__ PreRay get_pre() const {
        const int t = threadIdx.x;
        return {s->pre[0][t], s->pre[1][t], s->pre[2][t], s->pre[3][t], s->pre[4][t], s->pre[5][t], s->pre[6][t]};
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
};

// Stackless, threaded traversal.  The host flattens the reference tree in PREORDER, so the reference's
// recursion (bvhwrapper.rs:97-126: box test, then left subtree, then right subtree, one running
// closest t) visits nodes in increasing index order and a failed box test jumps to the node's SKIP
// link (first node after its subtree).  No stack, no depth limit:
//     i = 0; while (i < n_nodes) { if (box(i) passes) { test leaf primitives; i = i + 1 } else i = skip(i) }
// Node words: inner node  a = skip link (+ BIGBOX_BIT),          b = split axis (unused here)
//             leaf node   a = left primitive ref (+ BIGBOX_BIT), b = right primitive ref or REF_NONE (skip = i + 1)
// (a node built from a span of 1 or 2 holds its primitives directly, a node built from a span >= 3
// has two node children: bvhwrapper.rs:57-74.)  Leaves are tested with no box of their own, left
// first, right with the updated interval, strict comparisons — the reference's order.
template <typename R, typename Store, bool ANIM>
struct Trav {
    NodeRay<R> nr;
    Store store;
    float best32;     // (float)best_t, refreshed whenever best_t changes
    uint32_t i;       // next node index
    uint32_t wa, wb;  // words of the node tested last; after ST_LEAF they are the parked leaf primitives
    bool ok;          // the ray is regular and representable in f32: the cheap node step may be used

    __device__ __forceinline__ int walk_state() const { return ok ? (int)ST_NODE : (int)ST_EXACT; }
    // Box tests of INNER nodes are pure culling (for regular rays).  IEEE subtraction, multiplication, min and max are
    // monotone, a child's box lies inside its parent's exactly (Aabb::new_from_boxes is an exact min / max) and the
    // running closest t only shrinks, so whenever the reference's Aabb::hit (bvh.rs:96-132) fails on a node it fails on
    // every node of its subtree, for the interval of that moment and for every later one.  The primitives the reference
    // tests are therefore exactly those of the LEAF nodes whose own box passes at the moment DFS order reaches them;
    // what happens at inner nodes only decides how much work is skipped.  Two consequences used below:
    //   - the root of a tree with more than one node is not tested at all (the walk starts at its left child);
    //   - an inner node the f32 filter cannot decide may be entered without the exact f64 test (step_node).
    // Neither holds for a scene with nested HitLists: their members are tested with no box of their own, so the box node
    // above them is the one that selects them.  Such scenes set DevScene::strict_boxes: the walk starts at the root (the
    // caller passes n_nodes = 1) and every undecided box gets the exact test (free_pass_nodes = 0, free_pass_k = inf).
    // Irregular rays (zero / NaN / inf components) keep the reference's test at every node.
    __device__ __forceinline__ void init_from(const FilterRay& f, R tmax, uint32_t n_nodes) {
        nr.set(f);
        ok = f.ok;
        store.set_pre(f);
        store.set_best(tmax, REF_MISS);
        best32 = (float)tmax;
        i = (f.ok && n_nodes > 1u) ? 1u : 0u;
        wa = wb = REF_NONE;
    }
    // the box test of node i (words wa, wb) has been decided
    template <bool KNOWN_OK>
    __device__ __forceinline__ int after_box(bool hit, uint32_t n_nodes) {
        const bool leafnode = ref_is_leaf(wa);
        i = (hit || leafnode) ? i + 1u : (wa & INDEX_MASK);
        if (hit && leafnode) return ST_LEAF;
        return i >= n_nodes ? (int)ST_DONE : (KNOWN_OK ? (int)ST_NODE : walk_state());
    }
    // NODE step (precondition: ok).  f64 path: conservative filter (may answer ST_EXACT); f32 path: the f32 box
    // test itself (regular rays: min/max form == comparison form; irregular rays use step_exact).
    __device__ __forceinline__ int step_node(const DevScene<R>& sc, R tmin) {
        const NodeRec<float> nf = ldg_node32(sc.nodes32 + i);
        wa = nf.left;
        wb = nf.right;
        bool hit;
        if constexpr (sizeof(R) == 8) {
            const int dec = filter_box(nf, nr, (float)tmin, best32);
            // only a leaf node's decision selects primitives; an undecided inner node with a small subtree is entered
            // untested (at most free_pass_nodes cheap tests instead of one exact f64 test; unbounded, rays with a wide
            // error band would walk whole subtrees the exact test culls: 8x slower on the 10 M-triangle scene)
            if (dec == 0) {
                if (wa & ALWAYS_PASS_BIT) return after_box<true>(true, sc.n_nodes);  // member of a nested HitList: no box test
                if (ref_is_leaf(wa)) return ST_EXACT;
                if ((wa & INDEX_MASK) - i > sc.free_pass_nodes &&
                    !box_wider_than_band(ldg_node32(sc.nodes32 + i), nr, (float)tmin, best32, sc.free_pass_k))
                    return ST_EXACT;
            }
            hit = dec >= 0;
        } else {
            float lo, hi;
            slab32(nf, nr, tmin, best32, lo, hi);
            hit = hi > lo || (wa & ALWAYS_PASS_BIT) != 0u;
        }
        return after_box<true>(hit, sc.n_nodes);
    }
    // exact box test of node i in R arithmetic
    __device__ __forceinline__ int step_exact(const DevScene<R>& sc, V3<R> o, V3<R> d, R tmin) {
        const V3<R> inv = {R(1) / d.x, R(1) / d.y, R(1) / d.z};
        const NodeRec<R> n = ldg_rec<sizeof(NodeRec<R>) / 16>(sc.nodes + i);
        wa = n.left;
        wb = n.right;
        const R best = store.best_t();
        const bool hit = (wa & ALWAYS_PASS_BIT) ? true
                                                : (ok ? aabb_hit_regular(n, o, inv, inv.x > R(0), inv.y > R(0), inv.z > R(0), tmin, best)
                                                      : aabb_hit(n, o, inv, tmin, best));
        return after_box<false>(hit, sc.n_nodes);
    }
    __device__ __forceinline__ uint32_t leaf_left() const { return wa & ~(BIGBOX_BIT | ALWAYS_PASS_BIT); }
    __device__ __forceinline__ uint32_t leaf_right() const { return wb; }
    static __device__ __forceinline__ bool test_prim(const DevScene<R>& sc, uint32_t ref, V3<R> o, V3<R> d, R a, R tmin, R best, R tm,
                                                     R& t) {
        const uint32_t kind = ref_kind(ref), idx = ref_index(ref);
        if (kind == CR_PRIM_SPHERE) {
            const SphereRec<R> s = sphere_at<R, ANIM>(sc, idx, tm);
            return sphere_hit_t(s, o, d, a, tmin, best, t);
        } else if (kind == CR_PRIM_TRIANGLE) {
            const TriRec<R> tr = tri_at<R, ANIM>(sc, idx, tm);
            return tri_hit_t(tr, o, d, tmin, best, t);
        } else {
            const QuadRec<R> q = ldg_rec<sizeof(QuadRec<R>) / 16>(sc.quads + idx);
            R al, be;
            return quad_hit_t(q, o, d, tmin, best, t, al, be);
        }
    }
    // f64 path: are all primitives of the parked leaf node certain misses (nothing to test)?
    __device__ __forceinline__ bool leaf_certain_miss(const DevScene<R>& sc) const {
        if constexpr (sizeof(R) == 8) {
            const uint32_t pl = leaf_left(), pr = leaf_right();
            if (!ok || ref_kind(pl) != CR_PRIM_SPHERE) return false;
            // the pre-filter reads the construction-time f32 sphere: an animated one is always tested in full
            if (sphere_is_animated<R, ANIM>(sc, ref_index(pl))) return false;
            if (pr != REF_NONE && ref_kind(pr) == CR_PRIM_SPHERE && sphere_is_animated<R, ANIM>(sc, ref_index(pr))) return false;
            const PreRay pre = store.get_pre();
            if (!sphere_definite_miss(ldg_rec<1>(sc.spheres32 + ref_index(pl)), pre)) return false;
            if (pr == REF_NONE) return true;
            return ref_kind(pr) == CR_PRIM_SPHERE && sphere_definite_miss(ldg_rec<1>(sc.spheres32 + ref_index(pr)), pre);
        } else {
            return false;
        }
    }
    __device__ __forceinline__ int after_leaf(const DevScene<R>& sc) const { return i >= sc.n_nodes ? (int)ST_DONE : walk_state(); }
    __device__ __forceinline__ int step_leaf(const DevScene<R>& sc, V3<R> o, V3<R> d, R tmin) {
        const R a = vlen2(d);  // sphere.rs:74
        R best = store.best_t(), t;
        R tm = R(0);
        if constexpr (ANIM) tm = store.time();
        uint32_t bref = REF_NONE;
        const uint32_t pl = leaf_left(), pr = leaf_right();
        if (test_prim(sc, pl, o, d, a, tmin, best, tm, t)) {
            best = t;
            bref = pl;
        }
        if (pr != REF_NONE && test_prim(sc, pr, o, d, a, tmin, best, tm, t)) {
            best = t;
            bref = pr;
        }
        if (bref != REF_NONE) {
            store.set_best(best, bref);
            best32 = (float)best;
        }
        return after_leaf(sc);
    }
};

// Warp-persistent engine.  The warp alternates between
//   - a NODE phase: up to node_slice cheap box tests per lane.  A lane whose node the filter cannot
//     decide parks in ST_EXACT, a lane that entered a leaf node with a possible hit parks in ST_LEAF;
//   - an EXACT phase: the exact box test for the parked undecided nodes (a few % of the tests);
//   - a LEAF phase: the primitive tests of the parked leaf nodes.
// Parking the rare, expensive steps lets them run with several lanes at once instead of dragging the
// whole warp along for one lane.  Finished lanes are refilled (one atomic per warp) as soon as REFILL
// of them are idle.  Each lane's own sequence of tests is the reference's, so is the result.
//   IO::count() / cursor() / filter(i,tmin,tmax) / load(i,o,d) / commit(has,i,ref,t,o,d)  (commit is warp-synchronous)
template <typename R, int REFILL, int BLOCK, bool ANIM, typename IO>
__device__ __forceinline__ void trace_persistent(const DevScene<R>& sc, R tmin, R tmax, IO& io, LaneSlots<R, BLOCK, ANIM>* slots) {
    const int NODE_SLICE = sc.node_slice;
    const uint32_t n = io.count();
    const int lane = threadIdx.x & 31;
    Trav<R, SmemStore<R, BLOCK, ANIM>, ANIM> tv;
    tv.store.s = slots;
    tv.ok = false;
    tv.i = 0;
    tv.wa = tv.wb = REF_NONE;
    tv.best32 = (float)tmax;
    int st = ST_IDLE;
    bool exhausted = false;
    const bool small = n <= gridDim.x * blockDim.x;
    for (;;) {
        const uint32_t walking = __ballot_sync(0xffffffffu, st == ST_NODE || st == ST_EXACT || st == ST_LEAF);
        const int n_free = 32 - __popc(walking);
        if ((!exhausted && n_free >= REFILL) || walking == 0u) {  // warp-uniform
            {
                V3<R> o = {R(0), R(0), R(0)}, d = {R(0), R(0), R(0)};
                uint32_t my = 0, bref = REF_MISS;
                R bt = tmax, tm = R(0);
                if (st == ST_DONE) {
                    my = tv.store.my();
                    bref = tv.store.best_ref();
                    bt = tv.store.best_t();
                    if (io.commit_needs_ray()) {
                        tv.store.get_ray(o, d);
                        if constexpr (ANIM) tm = tv.store.time();
                    }
                }
                io.commit(st == ST_DONE, my, bref, bt, o, d, tm);
            }
            if (st == ST_DONE) st = ST_IDLE;
            if (!exhausted) {
                uint32_t base = 0;
                if (small) {
                    // few rays (the tail of a render): one static batch per warp, no atomic.  With thousands
                    // of resident warps the shared cursor alone costs ~70 us per launch (ncu launch list).
                    base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u;
                } else {
                    if (lane == 0) base = atomicAdd(io.cursor(), (uint32_t)n_free);
                    base = __shfl_sync(0xffffffffu, base, 0);
                }
                if (st == ST_IDLE) {
                    const uint32_t k = base + (uint32_t)__popc(~walking & ((1u << lane) - 1u));
                    if (k < n) {
                        tv.store.set_my(k);
                        V3<R> o, d;
                        io.load(k, o, d);  // the R-precision ray of the exact / leaf steps goes to the lane's slot
                        tv.init_from(io.filter(k, tmin, tmax), tmax, sc.strict_boxes ? 1u : sc.n_nodes);
                        tv.store.set_ray(o, d);
                        if constexpr (ANIM) tv.store.set_time(io.time(k));
                        st = (sc.n_nodes == 0u) ? (int)ST_DONE : tv.walk_state();
                    }
                }
                if (small || base + (uint32_t)n_free >= n) exhausted = true;
            }
            if (__ballot_sync(0xffffffffu, st != ST_IDLE) == 0u) break;  // nothing walking, nothing pending
        }
        // ---- NODE phase
        // (the slice ends early once fewer than MIN_NODE_LANES lanes still have a cheap step to do)
#pragma unroll 1
        for (int k = 0; k < NODE_SLICE; k += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (st == ST_NODE) st = tv.step_node(sc, tmin);
            }
            if (__popc(__ballot_sync(0xffffffffu, st == ST_NODE)) < sc.min_node_lanes) break;
        }
        // ---- LEAF pre-filter: leaf nodes whose primitives are all certain misses need no f64 work
        if (st == ST_LEAF && tv.leaf_certain_miss(sc)) st = tv.after_leaf(sc);
        // ---- EXACT + LEAF phases share one read of the lane's ray
        if (__any_sync(0xffffffffu, st == ST_EXACT || st == ST_LEAF)) {
            if (st == ST_EXACT || st == ST_LEAF) {
                V3<R> o, d;
                tv.store.get_ray(o, d);
                if (st == ST_EXACT) st = tv.step_exact(sc, o, d, tmin);
                if (st == ST_LEAF) st = tv.step_leaf(sc, o, d, tmin);
            }
        }
    }
}

// Closest hit for ONE ray per lane, all 32 lanes of the warp together (no refill, no queues): the same NODE
// slices and EXACT + LEAF phases as trace_persistent, so the rare expensive steps still run with several lanes
// at once.  Used by the tail kernel, where a warp owns 32 paths from their current segment to their end.
// Lanes with active == false only take part in the votes.  Returns the closest hit in (ref, t).
template <typename R, int BLOCK, bool ANIM>
__device__ __forceinline__ void trace_warp_batch(const DevScene<R>& sc, R tmin, R tmax, bool active, V3<R> o, V3<R> d, R tm,
                                                 LaneSlots<R, BLOCK, ANIM>* slots, uint32_t& ref, R& t) {
    Trav<R, SmemStore<R, BLOCK, ANIM>, ANIM> tv;
    tv.store.s = slots;
    tv.ok = false;
    tv.i = 0;
    tv.wa = tv.wb = REF_NONE;
    tv.best32 = (float)tmax;
    int st = ST_DONE;
    if (active) {
        const V3<R> inv = {R(1) / d.x, R(1) / d.y, R(1) / d.z};  // adinv, bvh.rs:111
        tv.init_from(make_filter_ray<R>(o, d, inv, tmin, tmax, sc.bsmall, sc.bmax), tmax, sc.strict_boxes ? 1u : sc.n_nodes);
        tv.store.set_ray(o, d);
        if constexpr (ANIM) tv.store.set_time(tm);
        st = (sc.n_nodes == 0u) ? (int)ST_DONE : tv.walk_state();
    }
    while (__any_sync(0xffffffffu, st != ST_DONE)) {
#pragma unroll 1
        for (int k = 0; k < sc.node_slice; k += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (st == ST_NODE) st = tv.step_node(sc, tmin);
            }
            if (__popc(__ballot_sync(0xffffffffu, st == ST_NODE)) < sc.min_node_lanes) break;
        }
        if (st == ST_LEAF && tv.leaf_certain_miss(sc)) st = tv.after_leaf(sc);
        if (st == ST_EXACT || st == ST_LEAF) {
            V3<R> ro, rd;
            tv.store.get_ray(ro, rd);
            if (st == ST_EXACT) st = tv.step_exact(sc, ro, rd, tmin);
            if (st == ST_LEAF) st = tv.step_leaf(sc, ro, rd, tmin);
        }
    }
    ref = REF_MISS;
    t = tmax;
    if (active) {
        ref = tv.store.best_ref();
        t = tv.store.best_t();
    }
}

}  // namespace crb
