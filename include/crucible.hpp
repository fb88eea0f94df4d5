// crucible.hpp — C++17 host-side mirror of the reference's scene / camera / material / texture API on top
// of the C ABI (crucible_gpu.h).  The reference is a Rust crate; no Rust toolchain exists in the build
// image, so the caller side of the boundary is provided in C++ with the reference's names, argument
// meaning and error behaviour (a Rust panic becomes a std::runtime_error).  Header only; link with
// -lcrucible_b200.  Citations are paths under the reference repository.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "crucible_gpu.h"

namespace crucible {

// ---- src/utils.rs
struct Point3 {
    double x = 0, y = 0, z = 0;
    Point3() = default;
    Point3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    static Point3 origin() { return {}; }
    Point3 operator-() const { return {-x, -y, -z}; }
    Point3 operator+(const Point3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Point3 operator-(const Point3& o) const { return *this + (-o); }  // utils.rs:295-301
    double length_squared() const { return x * x + y * y + z * z; }
    double length() const { return std::sqrt(length_squared()); }
};
inline Point3 operator*(double s, const Point3& v) { return {s * v.x, s * v.y, s * v.z}; }
using Vec3 = Point3;

struct Color {  // utils.rs:340-356: panics outside [0,1]
    double r, g, b;
    Color(double r_, double g_, double b_) : r(r_), g(g_), b(b_) {
        auto chk = [](double v, const char* n) {
            if (!(v <= 1.0)) throw std::runtime_error(std::string(n) + " must be lower than 1.0. Got " + std::to_string(v));
            if (!(v >= 0.0)) throw std::runtime_error(std::string(n) + " must be greater or equal to 0.0. Got " + std::to_string(v));
        };
        chk(r, "R"); chk(g, "G"); chk(b, "B");
    }
    Color operator*(const Color& o) const {  // utils.rs:576-590 (clamped)
        auto c = [](double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); };
        return {c(r * o.r), c(g * o.g), c(b * o.b)};
    }
};

// ---- src/textures
struct Textures {
    int kind = CR_TEX_SOLID;
    Color albedo{0, 0, 0};
    double inv_scale = 1.0;
    std::shared_ptr<Textures> even, odd;
    std::vector<uint8_t> rgb8;
    int w = 0, h = 0;
    static std::shared_ptr<Textures> SolidColor(Color c) {
        auto t = std::make_shared<Textures>();
        t->albedo = c;
        return t;
    }
    static std::shared_ptr<Textures> CheckerTexture(double scale, Color c1, Color c2) {  // checker_texture.rs:30-37
        auto t = std::make_shared<Textures>();
        t->kind = CR_TEX_CHECKER;
        t->inv_scale = 1.0 / scale;
        t->even = SolidColor(c1);
        t->odd = SolidColor(c2);
        return t;
    }
    static std::shared_ptr<Textures> ImageTexture(std::vector<uint8_t> rgb, int w_, int h_) {  // decoded RGB8 (img_loader.rs:28)
        auto t = std::make_shared<Textures>();
        t->kind = CR_TEX_IMAGE;
        t->rgb8 = std::move(rgb);
        t->w = w_;
        t->h = h_;
        return t;
    }
};

// ---- src/materials
struct Materials {
    int kind = CR_MAT_LAMBERTIAN;
    std::shared_ptr<Textures> tex;
    double scatter_prob = 1.0, fuzz = 0.0, ior = 1.5;
    Color albedo{0, 0, 0};
    double emit[3] = {0, 0, 0};
    static std::shared_ptr<Materials> Lambertian(Color c, double prob) { return LambertianTex(Textures::SolidColor(c), prob); }
    static std::shared_ptr<Materials> LambertianTex(std::shared_ptr<Textures> t, double prob) {
        auto m = std::make_shared<Materials>();
        m->tex = std::move(t);
        m->scatter_prob = prob;
        return m;
    }
    static std::shared_ptr<Materials> Metal(Color c, double fuzz) {  // metal.rs:16-27
        if (!(fuzz <= 1.0)) throw std::runtime_error("A metal cannot have a fuzz factor above 1.0");
        if (!(fuzz >= 0.0)) throw std::runtime_error("A metal cannot have a fuzz factor below 0.0");
        auto m = std::make_shared<Materials>();
        m->kind = CR_MAT_METAL;
        m->albedo = c;
        m->fuzz = fuzz;
        return m;
    }
    static std::shared_ptr<Materials> Dielectric(double refraction_index) {
        auto m = std::make_shared<Materials>();
        m->kind = CR_MAT_DIELECTRIC;
        m->ior = refraction_index;
        return m;
    }
    static std::shared_ptr<Materials> Emissive(double r, double g, double b) {  // EXTENSION
        auto m = std::make_shared<Materials>();
        m->kind = CR_MAT_EMISSIVE;
        m->emit[0] = r; m->emit[1] = g; m->emit[2] = b;
        return m;
    }
};

// ---- src/objects
struct Sphere {
    Point3 center;
    double radius;
    std::shared_ptr<Materials> mat;
    Sphere(Point3 c, double r, std::shared_ptr<Materials> m) : center(c), radius(r), mat(std::move(m)) {
        if (!(r >= 0.0)) throw std::runtime_error("Cannot make a sphere with negative radius");  // sphere.rs:26
    }
};
struct Triangle {
    Point3 a, b, c;
    std::shared_ptr<Materials> mat;
};
struct Quad {  // EXTENSION
    Point3 q, u, v;
    std::shared_ptr<Materials> mat;
};

enum class InterpolationType { NERP = CR_NERP, LERP = CR_LERP };
enum class TransformSpace { World, Local };

// ---- src/timeline (translate keyframes of a point), transform_builder.rs:348-727
class TransformTimeline {
  public:
    explicit TransformTimeline(Point3 start = {}) : start_(start) {}
    void translate_point(Point3 p, double keyframe, InterpolationType it, TransformSpace sp) {
        translate(0, p.x, keyframe, it, sp);
        translate(1, p.y, keyframe, it, sp);
        translate(2, p.z, keyframe, it, sp);
    }
    void translate(int axis, double x, double keyframe, InterpolationType it, TransformSpace sp) {
        if (!(keyframe >= 0.0)) throw std::runtime_error("Cannot add a keyframe before the animation start.");
        const double init[3] = {start_.x, start_.y, start_.z};
        double prev_end = init[axis], prev_time = 0.0;
        for (auto r = keys_.rbegin(); r != keys_.rend(); ++r)
            if (keyframe > r->k.t1 && r->k.axis == axis) {
                prev_end = r->end;
                prev_time = r->k.t1 > 0.0 ? r->k.t1 : 0.0;
                break;
            }
        Key k;
        k.end = x;
        k.k.delta = sp == TransformSpace::World ? x - prev_end : x;
        k.k.axis = axis;
        k.k.interp = (int)it;
        k.k.t0 = it == InterpolationType::LERP ? prev_time : keyframe;
        k.k.t1 = keyframe;
        keys_.push_back(k);
        std::stable_sort(keys_.begin(), keys_.end(), [](const Key& a, const Key& b) { return a.k.t0 < b.k.t0; });
    }
    Point3 start() const { return start_; }
    uint32_t fill(CrKeyframe* out) const {
        if (keys_.size() > CR_MAX_CAM_KEYS) throw std::runtime_error("too many camera keyframes");
        for (size_t i = 0; i < keys_.size(); ++i) out[i] = keys_[i].k;
        return (uint32_t)keys_.size();
    }
    Point3 combine_and_compute(double t) const {  // timeline/mod.rs:233-263
        std::vector<CrKeyframe> k(keys_.size());
        for (size_t i = 0; i < keys_.size(); ++i) k[i] = keys_[i].k;
        const double init[3] = {start_.x, start_.y, start_.z};
        double out[3];
        if (cr_camera_point_at(init, k.data(), k.size(), t, out)) throw std::runtime_error(cr_last_error());
        return {out[0], out[1], out[2]};
    }

  private:
    struct Key {
        CrKeyframe k;
        double end;
    };
    Point3 start_;
    std::vector<Key> keys_;
};

enum class Backend { Gpu };  // the CPU thread pool of the reference is what this library replaces: no fallback

// ---- src/camera/mod.rs:66-268
class Camera {
  public:
    Camera(double aspect_ratio, uint32_t image_width, double frame_rate, double shutter_angle, size_t thread_count = 0)
        : image_width_(image_width), frame_rate_(frame_rate), shutter_angle_(shutter_angle), threads_(thread_count) {
        const uint32_t h = (uint32_t)((double)image_width / aspect_ratio);  // Viewport::new, :36-47
        image_height_ = h < 1 ? 1 : h;
        fix_viewport();
    }
    void next_frame() { ++frame_; }
    void look_from(Point3 p) { from_ = TransformTimeline(p); fix_viewport(); }
    void look_at(Point3 p) { at_ = TransformTimeline(p); fix_viewport(); }
    void set_vup(Vec3 v) { vup_ = v; }
    void set_vfov(double deg) { vfov_ = deg * M_PI / 180.0; fix_viewport(); }
    void set_samples(uint32_t s) {
        if (s == 0) throw std::runtime_error("The camera must have a positive number of samples. 0 is invalid.");
        samples_ = s;
    }
    void set_max_depth(uint32_t d) { max_depth_ = d; }
    void set_defocus_angle(double deg) { defocus_angle_ = deg * M_PI / 180.0; }
    void set_focus_dist(double fd) { focus_dist_ = fd; fix_viewport(); }
    void set_threads(size_t t) { threads_ = t; }
    void set_device(int device) { device_ = device; }
    void set_seed(uint64_t seed) { seed_ = seed; }
    TransformTimeline& look_from_timeline() { return from_; }
    TransformTimeline& look_at_timeline() { return at_; }
    uint32_t width() const { return image_width_; }
    uint32_t height() const { return image_height_; }
    int device() const { return device_; }
    uint64_t seed() const { return seed_; }

    CrCamera to_abi() const {
        CrCamera c{};
        c.image_width = image_width_;
        c.image_height = image_height_;
        c.viewport_width = viewport_width_;
        c.viewport_height = viewport_height_;
        c.focus_dist = focus_dist_;
        c.defocus_angle = defocus_angle_;
        c.defocus_radius = focus_dist_ * std::tan(defocus_angle_ / 2.0);  // rendering_compute.rs:72-74
        const Point3 f = from_.start(), a = at_.start();
        c.vup[0] = vup_.x; c.vup[1] = vup_.y; c.vup[2] = vup_.z;
        c.look_from[0] = f.x; c.look_from[1] = f.y; c.look_from[2] = f.z;
        c.look_at[0] = a.x; c.look_at[1] = a.y; c.look_at[2] = a.z;
        c.frame_rate = frame_rate_;
        c.frame = (uint32_t)frame_;
        c.samples = samples_;
        c.max_depth = max_depth_;
        c.shutter_angle = shutter_angle_;
        c.n_from_keys = from_.fill(c.from_keys);
        c.n_at_keys = at_.fill(c.at_keys);
        return c;
    }

  private:
    void fix_viewport() {  // rendering_compute.rs:5-11
        const double h = std::tan(vfov_ / 2.0);
        viewport_height_ = 2.0 * h * focus_dist_;
        viewport_width_ = viewport_height_ * ((double)image_width_ / (double)image_height_);
    }
    uint32_t image_width_, image_height_ = 1;
    double vfov_ = 90.0 * M_PI / 180.0, viewport_width_ = 0, viewport_height_ = 0;
    TransformTimeline from_, at_;
    Vec3 vup_{0, 1, 0};
    double defocus_angle_ = 0.0, focus_dist_ = 10.0;
    uint32_t samples_ = 10, max_depth_ = 10;
    double frame_rate_, shutter_angle_;
    size_t frame_ = 0, threads_;
    int device_ = 0;
    uint64_t seed_ = 1;
};

// ---- src/scene/mod.rs:75-347
class Scene {
  public:
    Camera scene_cam;
    Scene(double aspect_ratio, uint32_t image_width, size_t frame_rate, double shutter_angle, size_t threads, double duration = -1.0)
        : scene_cam(aspect_ratio, image_width, (double)frame_rate, shutter_angle, threads), frame_rate_(frame_rate), duration_(duration) {}
    static Scene new_image(double ar, uint32_t w, size_t fr, double sa, size_t th) { return Scene(ar, w, fr, sa, th); }
    static Scene new_movie(double ar, uint32_t w, size_t fr, double sa, size_t th, double dur) { return Scene(ar, w, fr, sa, th, dur); }

    void add_element(const Sphere& s, const std::string& alias) {
        const int id = vend_id(alias);
        const double d[4] = {s.center.x, s.center.y, s.center.z, s.radius};
        push(CR_PRIM_SPHERE, d, 4, s.mat, id);
    }
    void add_element(const Triangle& t, const std::string& alias) {
        const int id = vend_id(alias);
        const double d[9] = {t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z};
        push(CR_PRIM_TRIANGLE, d, 9, t.mat, id);
    }
    void add_element(const Quad& q, const std::string& alias) {
        const int id = vend_id(alias);
        const double d[9] = {q.q.x, q.q.y, q.q.z, q.u.x, q.u.y, q.u.z, q.v.x, q.v.y, q.v.z};
        push(CR_PRIM_QUAD, d, 9, q.mat, id);
    }
    // Scene::load_asset (scene/mod.rs:191-230) with the reference's OBJ rules (obj_loader.rs:64-143)
    void load_asset(const std::string& path, const std::string& alias, double scale, Point3 shift, std::shared_ptr<Materials> mat) {
        const int id = vend_id(alias);
        std::ifstream in(path);
        if (!in) throw std::runtime_error("Cannot open OBJ file.");
        std::vector<Point3> v;
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ss(line);
            std::string tag;
            if (!(ss >> tag)) continue;
            if (tag == "v") {
                double x, y, z;
                if (!(ss >> x >> y >> z)) throw std::runtime_error("Invalid OBJ file. Expected a floating point value for a vertex.");
                v.push_back(scale * Point3(x, y, z) + shift);
            } else if (tag == "f") {
                size_t a, b, c;
                std::string extra;
                if (!(ss >> a >> b >> c) || (ss >> extra))
                    throw std::runtime_error("The asset loader only supports triangularized images, please triangulate the image then try again");
                const Point3 &A = v.at(a - 1), &B = v.at(b - 1), &C = v.at(c - 1);
                const double d[9] = {A.x, A.y, A.z, B.x, B.y, B.z, C.x, C.y, C.z};
                push(CR_PRIM_TRIANGLE, d, 9, mat, id);
            } else {
                throw std::runtime_error("Unsupported OBJ file");
            }
        }
    }
    void hide_element(const std::string& alias) { set_visibility(alias, true); }
    void show_element(const std::string& alias) { set_visibility(alias, false); }
    void load_default_skybox() { sky_kind_ = CR_SKY_DEFAULT; }
    void load_black_skybox() { sky_kind_ = CR_SKY_BLACK; }  // EXTENSION
    void load_spherical_skybox(std::vector<uint8_t> rgb8, int w, int h) {
        sky_kind_ = CR_SKY_SPHERICAL;
        sky_ = Textures::ImageTexture(std::move(rgb8), w, h);
    }
    void cam_translate_point(Point3 p, double keyframe, InterpolationType it, TransformSpace sp, const std::string& which) {
        (which == "from" ? scene_cam.look_from_timeline() : scene_cam.look_at_timeline()).translate_point(p, keyframe, it, sp);
    }
    size_t compute_frame_count() const { return (size_t)std::ceil(duration_ * (double)frame_rate_); }  // scene/mod.rs:324-330
    size_t n_prims() const { return prim_kind_.size(); }

    // Flatten into a CrScene (what a `mod gpu` inside the crate would do with its private fields)
    CrScene* flatten(int device) const {
        CrScene* s = cr_scene_create(device);
        if (!s) throw std::runtime_error(cr_last_error());
        try {
            // tables
            std::vector<CrMaterial> mats;
            std::vector<CrTexture> texs;
            std::map<const Materials*, int> mat_id;
            std::map<const Textures*, int> tex_id;
            std::vector<const Textures*> images;
            std::function<int(const std::shared_ptr<Textures>&)> intern_tex = [&](const std::shared_ptr<Textures>& t) -> int {
                auto it = tex_id.find(t.get());
                if (it != tex_id.end()) return it->second;
                CrTexture c{};
                c.kind = t->kind;
                if (t->kind == CR_TEX_SOLID) {
                    c.color[0] = t->albedo.r; c.color[1] = t->albedo.g; c.color[2] = t->albedo.b;
                } else if (t->kind == CR_TEX_CHECKER) {
                    c.inv_scale = t->inv_scale;
                    c.even = intern_tex(t->even);
                    c.odd = intern_tex(t->odd);
                } else {
                    c.image = cr_scene_add_image(s, t->rgb8.data(), t->w, t->h);
                    if (c.image < 0) throw std::runtime_error(cr_last_error());
                }
                texs.push_back(c);
                return tex_id[t.get()] = (int)texs.size() - 1;
            };
            std::vector<int32_t> prim_mat(prim_kind_.size());
            for (size_t i = 0; i < prim_kind_.size(); ++i) {
                const Materials* m = prim_mat_[i].get();
                auto it = mat_id.find(m);
                if (it == mat_id.end()) {
                    CrMaterial c{};
                    c.kind = m->kind;
                    c.scatter_prob = m->scatter_prob;
                    c.fuzz = m->fuzz;
                    c.ior = m->ior;
                    c.albedo[0] = m->albedo.r; c.albedo[1] = m->albedo.g; c.albedo[2] = m->albedo.b;
                    for (int k = 0; k < 3; ++k) c.emit[k] = m->emit[k];
                    if (m->kind == CR_MAT_LAMBERTIAN) c.tex = intern_tex(prim_mat_[i]->tex);
                    mats.push_back(c);
                    it = mat_id.emplace(m, (int)mats.size() - 1).first;
                }
                prim_mat[i] = it->second;
            }
            // primitives in insertion order (runs of equal kind go in one call)
            size_t i = 0, off[3] = {0, 0, 0};
            while (i < prim_kind_.size()) {
                size_t j = i;
                while (j < prim_kind_.size() && prim_kind_[j] == prim_kind_[i]) ++j;
                const int kind = prim_kind_[i];
                const size_t stride = kind == CR_PRIM_SPHERE ? 4 : 9;
                const double* data = prim_data_[kind].data() + off[kind];
                int64_t rc = kind == CR_PRIM_SPHERE ? cr_scene_add_spheres(s, data, &prim_mat[i], &prim_id_[i], j - i)
                             : kind == CR_PRIM_TRIANGLE ? cr_scene_add_triangles(s, data, &prim_mat[i], &prim_id_[i], j - i)
                                                        : cr_scene_add_quads(s, data, &prim_mat[i], &prim_id_[i], j - i);
                if (rc < 0) throw std::runtime_error(cr_last_error());
                off[kind] += (j - i) * stride;
                i = j;
            }
            for (size_t p = 0; p < prim_id_.size(); ++p)
                if (hidden_.count(prim_id_[p]) && cr_scene_set_hidden(s, p, 1)) throw std::runtime_error(cr_last_error());
            int sky_image = -1;
            if (sky_kind_ == CR_SKY_SPHERICAL) sky_image = cr_scene_add_image(s, sky_->rgb8.data(), sky_->w, sky_->h);
            if (cr_scene_set_materials(s, mats.data(), mats.size()) || cr_scene_set_textures(s, texs.data(), texs.size()) ||
                cr_scene_set_sky(s, sky_kind_, sky_image) || cr_scene_commit(s))
                throw std::runtime_error(cr_last_error());
        } catch (...) {
            cr_scene_destroy(s);
            throw;
        }
        return s;
    }

    // Scene::render_scene -> render_image -> Camera::render (scene/mod.rs:283-347, camera/mod.rs:270-317):
    // rebuilds the BVH, renders on the GPU backend, writes "<fname>.ppm" (P3, one "r g b" line per pixel) --
    // the library's cr_render_to_file does the sample loop, the copy back and the file.
    CrStats render_scene(const std::string& fname) {
        CrScene* s = flatten(scene_cam.device());
        const CrCamera cam = scene_cam.to_abi();
        CrRenderOpts opts{};
        opts.seed = scene_cam.seed();
        CrStats st{};
        const int rc = cr_render_to_file(s, &cam, &opts, (fname + ".ppm").c_str(), CR_PPM_P3, &st);
        const std::string err = rc ? cr_last_error() : "";
        cr_scene_destroy(s);
        if (rc) throw std::runtime_error("Render failed. " + err);
        std::fprintf(stderr, "Successful render! Image stored at: %s.ppm\n", fname.c_str());
        return st;
    }

    // Scene::render_movie's frame loop (scene/mod.rs:295-322) for a static world: `frames` images into `dir`
    // ("image{frame:0>digits}.ppm"), frame k+1 traced while frame k is copied, formatted and written.
    // rank / world shard whole frames over GPUs.  The caller creates `dir` and runs ffmpeg as the reference does.
    std::vector<CrStats> render_frames(const std::string& dir, uint32_t frames, uint32_t rank = 0, uint32_t world = 1,
                                       int format = CR_PPM_P3) {
        CrScene* s = flatten(scene_cam.device());
        const CrCamera cam = scene_cam.to_abi();
        CrRenderOpts opts{};
        opts.seed = scene_cam.seed();
        std::vector<CrStats> st((frames > rank ? (frames - rank + world - 1) / world : 0) + 1);
        const uint32_t digits = (uint32_t)std::to_string(frames).size();
        const int rc = cr_render_frames(s, &cam, &opts, rank, world, frames, dir.c_str(), digits, format, st.data());
        const std::string err = rc ? cr_last_error() : "";
        cr_scene_destroy(s);
        if (rc) throw std::runtime_error("Render failed. " + err);
        st.pop_back();
        for (uint32_t f = 0; f < frames; ++f) scene_cam.next_frame();
        return st;
    }

  private:
    int vend_id(const std::string& alias) {  // scene/id_vendor.rs:28-42; a collision panics in the scene
        if (ids_.count(alias)) throw std::runtime_error("This alias collides with another name in the scene! Try changing " + alias + " to a new name.");
        const int id = (int)ids_.size();
        ids_[alias] = id;
        return id;
    }
    void push(int kind, const double* d, size_t n, const std::shared_ptr<Materials>& m, int id) {
        prim_kind_.push_back(kind);
        prim_data_[kind].insert(prim_data_[kind].end(), d, d + n);
        prim_mat_.push_back(m);
        prim_id_.push_back(id);
    }
    void set_visibility(const std::string& alias, bool hide) {
        auto it = ids_.find(alias);
        if (it == ids_.end()) {
            std::fprintf(stderr, "WARNING: The element `%s` does not exist. Are you sure you typed the right name?\n", alias.c_str());
            return;
        }
        if (hide) hidden_[it->second] = true; else hidden_.erase(it->second);
    }
    size_t frame_rate_;
    double duration_;
    std::vector<int> prim_kind_;
    std::vector<double> prim_data_[3];
    std::vector<std::shared_ptr<Materials>> prim_mat_;
    std::vector<int32_t> prim_id_;
    std::map<std::string, int> ids_;
    std::map<int, bool> hidden_;
    int sky_kind_ = CR_SKY_DEFAULT;
    std::shared_ptr<Textures> sky_;
};

// ---- src/demo_builder/demo_images.rs:14-109 with a SEEDED generator (the reference's is unseeded, :45)
namespace demo_images {
inline Scene book1_end_scene(size_t threads, uint64_t seed = 1, uint32_t image_width = 400, uint32_t samples = 500) {
    Scene sc = Scene::new_image(16.0 / 9.0, image_width, 24, 180.0, threads);
    sc.scene_cam.set_samples(samples);
    sc.scene_cam.set_max_depth(50);
    sc.scene_cam.look_from({13.0, 2.0, 3.0});
    sc.scene_cam.look_at({0.0, 0.0, 0.0});
    sc.scene_cam.set_vfov(20.0);
    sc.scene_cam.set_defocus_angle(0.6);
    sc.scene_cam.set_focus_dist(10.0);
    sc.scene_cam.set_seed(seed);
    auto checker = Textures::CheckerTexture(0.32, Color(0.2, 0.3, 0.1), Color(0.9, 0.9, 0.9));
    sc.add_element(Sphere({0.0, -1000.0, 0.0}, 1000.0, Materials::LambertianTex(checker, 1.0)), "ground");
    uint32_t ctr[4] = {0, 0, 0, 0}, buf[4];
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int have = 0;
    auto rnd = [&]() {  // Philox4x32-10 stream through the library's own block function
        if (have == 0) {
            cr_philox4x32_10(ctr, key, buf);
            ++ctr[0];
            have = 2;
        }
        const int h = 2 - have--;
        const uint64_t x = ((uint64_t)buf[2 * h + 1] << 32) | buf[2 * h];
        return (double)(x >> 11) * (1.0 / 9007199254740992.0);
    };
    int counter = 0;
    for (int a = -11; a < 11; ++a)
        for (int b = -11; b < 11; ++b) {
            const double choose_mat = rnd();
            const Point3 center(a + 0.9 * rnd(), 0.2, b + 0.9 * rnd());
            if ((center - Point3(4.0, 0.2, 0.0)).length() > 0.9) {
                std::shared_ptr<Materials> m;
                if (choose_mat < 0.8) {
                    const Color c1(rnd(), rnd(), rnd()), c2(rnd(), rnd(), rnd());
                    m = Materials::Lambertian(c1 * c2, 1.0);
                } else if (choose_mat < 0.95) {
                    const Color alb(0.5 + 0.5 * rnd(), 0.5 + 0.5 * rnd(), 0.5 + 0.5 * rnd());
                    m = Materials::Metal(alb, 0.5 * rnd());
                } else {
                    m = Materials::Dielectric(1.5);
                }
                sc.add_element(Sphere(center, 0.2, m), "small" + std::to_string(counter++));
            }
        }
    sc.add_element(Sphere({0.0, 1.0, 0.0}, 1.0, Materials::Dielectric(1.5)), "large_dielectric");
    sc.add_element(Sphere({-4.0, 1.0, 0.0}, 1.0, Materials::Lambertian(Color(0.4, 0.2, 0.1), 1.0)), "large_lambertian");
    sc.add_element(Sphere({4.0, 1.0, 0.0}, 1.0, Materials::Metal(Color(0.7, 0.6, 0.5), 0.0)), "large_metal");
    return sc;
}
}  // namespace demo_images

}  // namespace crucible
