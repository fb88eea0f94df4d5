"""Host-side mirror of the reference's scene / camera / material / texture / timeline API.

The reference is a Rust crate whose public surface is `scene`, `timeline`, `utils`, `demo_builder`
(src/lib.rs:7-10).  There is no rustc in this environment, so the caller of the C ABI is mirrored here
with the reference's names, argument meaning and error behaviour (panics become exceptions), so that
tests read like the reference's own code.  Everything below only *describes* a scene; the hot path
(`Camera::render`) is the CUDA library behind crucible_b200.gpu.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import abi


# ------------------------------------------------------------------ utils.rs
class Color(tuple):
    """Color::new (src/utils.rs:345-356): panics outside [0,1]."""

    def __new__(cls, r, g, b):
        for n, v in (("R", r), ("G", g), ("B", b)):
            if not (v <= 1.0):
                raise ValueError(f"{n} must be lower than 1.0. Got {v}")
            if not (v >= 0.0):
                raise ValueError(f"{n} must be greater or equal to 0.0. Got {v}")
        return super().__new__(cls, (float(r), float(g), float(b)))

    def __mul__(self, o):  # Color * Color, utils.rs:576-590
        return Color(*[min(max(a * b, 0.0), 1.0) for a, b in zip(self, o)])


def Point3(x, y, z):
    return np.array([x, y, z], dtype=np.float64)


Vec3 = Point3


# ------------------------------------------------------------------ textures
class SolidColor:  # src/textures/solid_color.rs
    def __init__(self, albedo: Color):
        self.albedo = Color(*albedo)

    @staticmethod
    def new_from_color(c):
        return SolidColor(c)

    @staticmethod
    def new_from_rgb(r, g, b):
        return SolidColor(Color(r, g, b))


class CheckerTexture:  # src/textures/checker_texture.rs:18-37
    def __init__(self, scale, even, odd):
        self.inv_scale = 1.0 / scale
        self.even, self.odd = even, odd

    @staticmethod
    def new_from_textures(scale, even, odd):
        return CheckerTexture(scale, even, odd)

    @staticmethod
    def new_from_color(scale, c1, c2):
        return CheckerTexture(scale, SolidColor(c1), SolidColor(c2))


class ImageTexture:  # src/textures/image_texture.rs; takes decoded RGB8 (the decode is host I/O)
    def __init__(self, rgb8: np.ndarray):
        rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
        if rgb8.ndim != 3 or rgb8.shape[2] != 3:
            raise ValueError("ImageTexture expects [h][w][3] uint8")
        self.rgb8 = rgb8


# ------------------------------------------------------------------ materials
class Lambertian:  # src/materials/lambertian.rs:24-38
    def __init__(self, tex, prob):
        self.tex, self.scatter_prob = tex, float(prob)

    @staticmethod
    def new_from_color(c, prob):
        return Lambertian(SolidColor(c), prob)

    @staticmethod
    def new_from_texture(tex, prob):
        return Lambertian(tex, prob)


class Metal:  # src/materials/metal.rs:16-27
    def __init__(self, c, fuzz):
        if not fuzz <= 1.0:
            raise ValueError("A metal cannot have a fuzz factor above 1.0")
        if not fuzz >= 0.0:
            raise ValueError("A metal cannot have a fuzz factor below 0.0")
        self.albedo, self.fuzz = Color(*c), float(fuzz)


class Dielectric:  # src/materials/dielectric.rs:13-19
    def __init__(self, refraction_index):
        self.refraction_index = float(refraction_index)


class Emissive:  # EXTENSION (absent in the reference)
    def __init__(self, rgb):
        self.emit = tuple(float(x) for x in rgb)


# ------------------------------------------------------------------ objects
class Sphere:  # src/objects/sphere.rs:24-39
    def __init__(self, center, radius, mat):
        if not radius >= 0.0:
            raise ValueError("Cannot make a sphere with negative radius")
        self.center, self.radius, self.mat = np.asarray(center, np.float64), float(radius), mat


class Triangle:  # src/objects/triangle.rs:22-46
    def __init__(self, a, b, c, mat):
        self.a, self.b, self.c = np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64)
        self.mat = mat


class Quad:  # EXTENSION
    def __init__(self, q, u, v, mat):
        self.q, self.u, self.v, self.mat = np.asarray(q, np.float64), np.asarray(u, np.float64), np.asarray(v, np.float64), mat


# ------------------------------------------------------------------ timeline (translate part)
class HitList:  # src/objects/hitlist.rs:6-31: a nested element, built with add()
    def __init__(self, objs=()):
        self.objs = []
        for o in objs:
            self.add(o)

    def add(self, obj):
        self.objs.append(obj)


class BVHWrapper:  # src/objects/bvhwrapper.rs:15-44
    def __init__(self, hitlist):
        self.list = hitlist

    @staticmethod
    def new_wrapper(hitlist):
        return BVHWrapper(hitlist)


class InterpolationType:
    NERP, LERP = abi.CR_NERP, abi.CR_LERP


class TransformSpace:
    World, Local = "World", "Local"


@dataclass
class _Transform:  # src/timeline/mod.rs:62-71
    t0: float
    t1: float
    axis: int  # -1 = Omni (the initial position)
    delta: float
    interp: int
    end: float  # TransformResult::Translate{X,Y,Z}(standard value)


class TransformTimeline:
    """Translate keyframes of src/timeline/{mod,transform_builder,helper_functions}.rs for a point."""

    def __init__(self, start_pos, start_scale=1.0):
        self.start_pos = np.asarray(start_pos, np.float64).copy()
        self.start_scale = float(start_scale)
        self.translate: list[_Transform] = []  # excludes the Omni init entry (valid_time (-0.1,-0.1))
        self.scale: list[_Transform] = []      # ScaleR keys: delta = start value, end = end value

    @staticmethod
    def new(start_pos, _start_rot=None, _start_scale=1.0):
        return TransformTimeline(start_pos, _start_scale)

    @staticmethod
    def new_sphere(start_pos, _start_rot, start_radius):  # timeline/mod.rs:178-223
        return TransformTimeline(start_pos, start_radius)

    def combine_and_compute_object(self, t):
        """timeline/mod.rs:233-263 for an object point: (x, y, z, w) with w = the radius / scale entry, through the
        library's host evaluator (cr_anim_point_at: the routine the kernels run at the ray's time)."""
        keys = self.anim_keys()
        arr = (abi.CrAnimKey * max(len(keys), 1))(*keys)
        init = (abi.C.c_double * 4)(*self.start_pos, self.start_scale)
        out = (abi.C.c_double * 4)()
        abi.check(abi.load().cr_anim_point_at(init, abi.C.cast(arr, abi.C.c_void_p), len(keys), float(t), out))
        return np.array(list(out))

    def _translate(self, axis, x, keyframe, interp, space):
        # transform_builder.rs:348-469
        if not keyframe >= 0.0:
            raise ValueError(f"Cannot add a keyframe before the animation start. You tried to add keyframe: {keyframe}")
        # most_recent_matching_transform (helper_functions.rs:42-130): last entry whose interval ends before
        # the keyframe and whose type is this axis or Omni
        prev_end, prev_time = float(self.start_pos[axis]), 0.0  # Omni init: max(-0.1, 0.0)
        for tf in reversed(self.translate):
            if keyframe > tf.t1 and tf.axis == axis:
                prev_end, prev_time = tf.end, max(tf.t1, 0.0)
                break
        standard = float(x)
        delta = standard - prev_end if space == TransformSpace.World else standard
        if interp == InterpolationType.LERP:
            t0, t1 = prev_time, float(keyframe)
        else:
            t0 = t1 = float(keyframe)
        self.translate.append(_Transform(t0, t1, axis, delta, interp, standard))
        # sort_by(compare_start): stable by interval start
        self.translate.sort(key=lambda tf: tf.t0)

    def translate_x(self, x, keyframe, interp, space):
        self._translate(0, x, keyframe, interp, space)

    def translate_y(self, y, keyframe, interp, space):
        self._translate(1, y, keyframe, interp, space)

    def translate_z(self, z, keyframe, interp, space):
        self._translate(2, z, keyframe, interp, space)

    def translate_point(self, p, keyframe, interp, space):  # transform_builder.rs:715-727
        self.translate_x(p[0], keyframe, interp, space)
        self.translate_y(p[1], keyframe, interp, space)
        self.translate_z(p[2], keyframe, interp, space)

    # ---- scale_sphere / scale_x / scale_y / scale_z, transform_builder.rs:18-346 (`start_scale` = the construction
    # radius of a sphere, 1.0 for Triangle::new's timelines).  kind 3 = ScaleR, 4 / 5 / 6 = ScaleX / ScaleY / ScaleZ.
    _SCALE_NAMES = {3: "r", 4: "x", 5: "y", 6: "z"}

    def _scale(self, kind, value, keyframe, interp):
        if not keyframe >= 0.0:
            raise ValueError("Cannot add a keyframe before the animation start. You tried to add keyframe: "
                             f"{keyframe} in a {self._SCALE_NAMES[kind]} scaling")
        # most_recent_matching_transform(keyframe, kind) (helper_functions.rs:42-93): last entry in LIST order whose
        # interval ended before the keyframe and whose type is this one or Omni (the init entry, valid_time (-0.1, -0.1))
        prev_end, prev_time = self.start_scale, 0.0
        for tf in reversed(self.scale):
            if keyframe > tf.t1 and tf.axis == kind:
                prev_end, prev_time = tf.end, max(tf.t1, 0.0)
                break
        if interp == InterpolationType.LERP:
            t0, t1 = prev_time, float(keyframe)
        else:
            t0 = t1 = float(keyframe)
        self.scale.append(_Transform(t0, t1, kind, float(prev_end), interp, float(value)))
        self.scale.sort(key=lambda tf: tf.t0)  # sort_by(compare_start): stable

    def scale_sphere(self, r, keyframe, interp):  # :18-96
        self._scale(3, r, keyframe, interp)

    def scale_x(self, x, keyframe, interp):  # :101-180
        self._scale(4, x, keyframe, interp)

    def scale_y(self, y, keyframe, interp):  # :186-265 (the matrix carries the value in row 1, column 0: :229-246)
        self._scale(5, y, keyframe, interp)

    def scale_z(self, z, keyframe, interp):  # :271-346
        self._scale(6, z, keyframe, interp)

    def scale_point(self, p, keyframe, interp):  # :729-733
        self.scale_x(p[0], keyframe, interp)
        self.scale_y(p[1], keyframe, interp)
        self.scale_z(p[2], keyframe, interp)

    def anim_keys(self):
        """The timeline beyond its init entries as CrAnimKey[] in evaluation order (translate list, then scale list)."""
        keys = [abi.CrAnimKey(tf.t0, tf.t1, tf.delta, 0.0, tf.axis, tf.interp) for tf in self.translate]
        keys += [abi.CrAnimKey(tf.t0, tf.t1, tf.delta, tf.end, tf.axis, tf.interp) for tf in self.scale]
        return keys

    def keyframes(self):
        out = (abi.CrKeyframe * abi.CR_MAX_CAM_KEYS)()
        if len(self.translate) > abi.CR_MAX_CAM_KEYS:
            raise ValueError("too many camera keyframes")
        for i, tf in enumerate(self.translate):
            out[i] = abi.CrKeyframe(tf.t0, tf.t1, tf.delta, tf.axis, tf.interp)
        return out, len(self.translate)

    def combine_and_compute(self, t):
        """timeline/mod.rs:233-263 through the library's host evaluator (x, y, z, w=1)."""
        keys, n = self.keyframes()
        out = (abi.C.c_double * 3)()
        init = (abi.C.c_double * 3)(*self.start_pos)
        abi.check(abi.load().cr_camera_point_at(init, abi.C.cast(keys, abi.C.c_void_p), n, float(t), out))
        return np.array([out[0], out[1], out[2], 1.0])


# ------------------------------------------------------------------ camera
class Camera:
    """src/camera/mod.rs:66-268 (state + setters).  `render` lives in crucible_b200.gpu."""

    def __init__(self, aspect_ratio, image_width, frame_rate, shutter_angle, thread_count=0):
        self.aspect_ratio = float(aspect_ratio)
        self.image_width = int(image_width)
        self.image_height = max(int(image_width / aspect_ratio), 1)  # Viewport::new, :36-47
        self.vfov = 90.0 * math.pi / 180.0  # Degrees::as_radians: d * PI / 180 (utils.rs:27-31)
        self.look_from_tl = TransformTimeline(Point3(0, 0, 0))
        self.look_at_tl = TransformTimeline(Point3(0, 0, 0))
        self.vup = Vec3(0.0, 1.0, 0.0)
        self.defocus_angle = 0.0
        self.focus_dist = 10.0
        self.samples = 10
        self.max_depth = 10
        self.thread_count = thread_count
        self.frame_rate = float(frame_rate)
        self.frame = 0
        self.shutter_angle = float(shutter_angle)
        # Camera::new computes the viewport from focal_length 1.0 (:119-122); every demo then calls a setter
        h = math.tan(self.vfov / 2.0)
        self.viewport_height = 2.0 * h * 1.0
        self.viewport_width = self.viewport_height * (self.image_width / self.image_height)

    def _fix_viewport(self):  # rendering_compute.rs:5-11
        h = math.tan(self.vfov / 2.0)
        self.viewport_height = 2.0 * h * self.focus_dist
        self.viewport_width = self.viewport_height * (self.image_width / self.image_height)

    def next_frame(self):
        self.frame += 1

    def look_from(self, loc):  # :187-195 (replaces the timeline)
        self.look_from_tl = TransformTimeline(loc)
        self._fix_viewport()

    def look_at(self, loc):  # :198-203
        self.look_at_tl = TransformTimeline(loc)
        self._fix_viewport()

    def set_vup(self, vup):
        self.vup = np.asarray(vup, np.float64)

    def set_vfov(self, deg):  # :211-215
        self.vfov = deg * math.pi / 180.0
        self._fix_viewport()

    def set_samples(self, s):  # :233-240
        if not s > 0:
            raise ValueError(f"The camera must have a positive number of samples. {s} is invalid.")
        self.samples = int(s)

    def set_max_depth(self, md):
        self.max_depth = int(md)

    def set_defocus_angle(self, deg):  # :249-251
        self.defocus_angle = deg * math.pi / 180.0

    def set_focus_dist(self, fd):  # :254-258
        self.focus_dist = float(fd)
        self._fix_viewport()

    def set_threads(self, threads):
        self.thread_count = threads

    def to_abi(self) -> abi.CrCamera:
        c = abi.CrCamera()
        c.image_width, c.image_height = self.image_width, self.image_height
        c.viewport_width, c.viewport_height = self.viewport_width, self.viewport_height
        c.focus_dist = self.focus_dist
        c.defocus_angle = self.defocus_angle
        c.defocus_radius = self.focus_dist * math.tan(self.defocus_angle / 2.0)  # rendering_compute.rs:72-74
        for k in range(3):
            c.vup[k] = float(self.vup[k])
            c.look_from[k] = float(self.look_from_tl.start_pos[k])
            c.look_at[k] = float(self.look_at_tl.start_pos[k])
        c.frame_rate, c.frame = self.frame_rate, self.frame
        c.samples, c.max_depth = self.samples, self.max_depth
        c.shutter_angle = self.shutter_angle
        fk, nf = self.look_from_tl.keyframes()
        ak, na = self.look_at_tl.keyframes()
        c.from_keys, c.n_from_keys = fk, nf
        c.at_keys, c.n_at_keys = ak, na
        return c


# ------------------------------------------------------------------ flat description handed to a backend
@dataclass
class SceneDesc:
    """Flat scene: primitive batches in insertion order + material / texture / image tables + sky."""
    batches: list = field(default_factory=list)  # (kind, data[n,k] f64, material[n] i32, obj_id[n] i32)
    materials: list = field(default_factory=list)  # abi.CrMaterial
    textures: list = field(default_factory=list)  # abi.CrTexture
    images: list = field(default_factory=list)  # np.uint8 [h][w][3]
    sky_kind: int = abi.CR_SKY_DEFAULT
    sky_image: int = -1
    hidden: list = field(default_factory=list)  # prim indices
    animation: list = field(default_factory=list)  # (prim_index, point, [abi.CrAnimKey]) per animated point

    @property
    def n_prims(self):
        return sum(len(b[1]) for b in self.batches if b[0] >= 0)

    def begin_group(self, kind):
        """A nested element (Hittables::HitList / BVHWrapper as one scene element): the batches up to the matching
        end_group() are its members."""
        self.batches.append((abi.GROUP_BEGIN, np.zeros((0, 1)), np.array([kind], np.int32), np.zeros(0, np.int32)))

    def end_group(self):
        self.batches.append((abi.GROUP_END, np.zeros((0, 1)), np.zeros(0, np.int32), np.zeros(0, np.int32)))

    def apply(self, lib, handle, prefix):
        """Push the description through a C ABI with the cr_scene_* shape (the product's, or the oracle's)."""
        C = abi.C
        f = lambda name: getattr(lib, prefix + name)
        add = {abi.CR_PRIM_SPHERE: "scene_add_spheres", abi.CR_PRIM_TRIANGLE: "scene_add_triangles",
               abi.CR_PRIM_QUAD: "scene_add_quads"}
        if hasattr(lib, prefix + "scene_reserve"):  # size hint: the staging arrays grow once (the oracle has no such call)
            counts = [sum(len(b[1]) for b in self.batches if b[0] == k) for k in (abi.CR_PRIM_SPHERE, abi.CR_PRIM_TRIANGLE, abi.CR_PRIM_QUAD)]
            rc = f("scene_reserve")(handle, *counts)
            if rc < 0:
                raise abi.CrucibleError(rc, f("last_error")().decode())
        cache = self.__dict__.setdefault("_c_args", {})
        live = {id(b[1]) for b in self.batches}
        for k in [k for k in cache if k not in live]:  # batches that were replaced since the last call
            del cache[k]

        def c_args(data, mat, oid):
            # the converted arrays and their pointers are kept with the batch: a mesh world is ~1 600 batches, and a caller
            # that rebuilds the scene every frame (Scene::render_image) would pay numpy / ctypes bookkeeping for each again
            hit = cache.get(id(data))
            if hit is None or hit[0] is not data:
                cd, cm, co = np.ascontiguousarray(data, np.float64), np.ascontiguousarray(mat, np.int32), np.ascontiguousarray(oid, np.int32)
                hit = (data, cd, cm, co, cd.ctypes.data_as(C.c_void_p), cm.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p))
                if cd is data and cm is mat and co is oid:  # only views of the batch's own memory are kept (no stale copies)
                    cache[id(data)] = hit
            return hit

        def flush(run):
            """The primitive batches collected since the last group marker: one cr_scene_add_batches call when the library
            has it (validated and copied on all host threads), else one add call per batch (the oracle)."""
            if not run:
                return
            if len(run) > 1 and hasattr(lib, prefix + "scene_add_batches"):
                n = len(run)
                hits = [c_args(d, m, o) for _, d, m, o in run]  # (keeps converted copies alive during the call)
                kinds = (C.c_int32 * n)(*[k for k, _, _, _ in run])
                counts = (C.c_size_t * n)(*[len(d) for _, d, _, _ in run])
                pd = (C.c_void_p * n)(*[h[4].value for h in hits])
                pm = (C.c_void_p * n)(*[h[5].value for h in hits])
                po = (C.c_void_p * n)(*[h[6].value for h in hits])
                rc = f("scene_add_batches")(handle, n, kinds, pd, pm, po, counts)
                if rc < 0:
                    raise abi.CrucibleError(rc, f("last_error")().decode())
            else:
                for kind, data, mat, oid in run:
                    hit = c_args(data, mat, oid)
                    rc = f(add[kind])(handle, hit[4], hit[5], hit[6], len(data))
                    if rc < 0:
                        raise abi.CrucibleError(rc, f(("last_error"))().decode())
            run.clear()

        run = []
        for kind, data, mat, oid in self.batches:
            if kind in (abi.GROUP_BEGIN, abi.GROUP_END):
                flush(run)
                rc = f("scene_begin_group")(handle, int(mat[0])) if kind == abi.GROUP_BEGIN else f("scene_end_group")(handle)
                if rc < 0:
                    raise abi.CrucibleError(rc, f("last_error")().decode())
                continue
            run.append((kind, data, mat, oid))
        flush(run)
        for im in self.images:
            im = np.ascontiguousarray(im, np.uint8)
            rc = f("scene_add_image")(handle, im.ctypes.data_as(C.c_void_p), im.shape[1], im.shape[0])
            if rc < 0:
                raise abi.CrucibleError(rc, f("last_error")().decode())
        mats = (abi.CrMaterial * max(len(self.materials), 1))(*self.materials)
        texs = (abi.CrTexture * max(len(self.textures), 1))(*self.textures)
        for rc in (f("scene_set_materials")(handle, C.cast(mats, C.c_void_p), len(self.materials)),
                   f("scene_set_textures")(handle, C.cast(texs, C.c_void_p), len(self.textures)),
                   f("scene_set_sky")(handle, self.sky_kind, self.sky_image)):
            if rc < 0:
                raise abi.CrucibleError(rc, f("last_error")().decode())
        for p in self.hidden:
            rc = f("scene_set_hidden")(handle, p, 1)
            if rc < 0:
                raise abi.CrucibleError(rc, f("last_error")().decode())
        for prim, point, keys in self.animation:
            arr = (abi.CrAnimKey * max(len(keys), 1))(*keys)
            rc = f("scene_set_keyframes")(handle, prim, point, C.cast(arr, C.c_void_p), len(keys))
            if rc < 0:
                raise abi.CrucibleError(rc, f("last_error")().decode())
        rc = f("scene_commit")(handle)
        if rc < 0:
            raise abi.CrucibleError(rc, f("last_error")().decode())


class _Tables:
    """Interns material / texture objects into the flat tables (identity based, insertion ordered)."""

    def __init__(self):
        self.materials, self.textures, self.images = [], [], []
        self._mat_ids, self._tex_ids = {}, {}

    def texture(self, t):
        if id(t) in self._tex_ids:
            return self._tex_ids[id(t)]
        ct = abi.CrTexture()
        if isinstance(t, SolidColor):
            ct.kind = abi.CR_TEX_SOLID
            for k in range(3):
                ct.color[k] = t.albedo[k]
        elif isinstance(t, CheckerTexture):
            ct.kind = abi.CR_TEX_CHECKER
            ct.inv_scale = t.inv_scale
            ct.even, ct.odd = self.texture(t.even), self.texture(t.odd)
        elif isinstance(t, ImageTexture):
            ct.kind = abi.CR_TEX_IMAGE
            self.images.append(t.rgb8)
            ct.image = len(self.images) - 1
        else:
            raise TypeError(f"unknown texture {t!r}")
        self.textures.append(ct)
        self._tex_ids[id(t)] = len(self.textures) - 1
        self._keep = getattr(self, "_keep", []) + [t]
        return len(self.textures) - 1

    def material(self, m):
        if id(m) in self._mat_ids:
            return self._mat_ids[id(m)]
        cm = abi.CrMaterial()
        cm.scatter_prob = 1.0
        if isinstance(m, Lambertian):
            cm.kind, cm.tex, cm.scatter_prob = abi.CR_MAT_LAMBERTIAN, self.texture(m.tex), m.scatter_prob
        elif isinstance(m, Metal):
            cm.kind, cm.fuzz = abi.CR_MAT_METAL, m.fuzz
            for k in range(3):
                cm.albedo[k] = m.albedo[k]
        elif isinstance(m, Dielectric):
            cm.kind, cm.ior = abi.CR_MAT_DIELECTRIC, m.refraction_index
        elif isinstance(m, Emissive):
            cm.kind = abi.CR_MAT_EMISSIVE
            for k in range(3):
                cm.emit[k] = m.emit[k]
        else:
            raise TypeError(f"unknown material {m!r}")
        self.materials.append(cm)
        self._mat_ids[id(m)] = len(self.materials) - 1
        self._keep = getattr(self, "_keep", []) + [m]
        return len(self.materials) - 1


class ObjectType:
    Camera, Sphere, TriangleMesh, Triangle, Quad = range(5)


class Scene:
    """src/scene/mod.rs:75-347: camera + flat element list + skybox.  `render_scene` selects the GPU backend."""

    def __init__(self, aspect_ratio, image_width, frame_rate, shutter_angle, thread_count=0, duration=None):
        self.scene_cam = Camera(aspect_ratio, image_width, float(frame_rate), shutter_angle, thread_count)
        self.frame_rate = frame_rate
        self.duration = duration
        self._tables = _Tables()
        self._batches = []  # [kind, data, mat, oid]
        self._aliases = {}  # IdVendor, scene/id_vendor.rs
        self._next_id = 0
        self._hidden_ids = set()
        self._anim_ops = {}  # object id -> [(op, args...)] in call order (scene_animator.rs)
        self.sky_kind, self._sky_rgb8 = abi.CR_SKY_DEFAULT, None

    @staticmethod
    def new_image(aspect_ratio, image_width, frame_rate, shutter_angle, thread_count=0):
        return Scene(aspect_ratio, image_width, frame_rate, shutter_angle, thread_count)

    @staticmethod
    def new_movie(aspect_ratio, image_width, frame_rate, shutter_angle, thread_count, duration):
        return Scene(aspect_ratio, image_width, frame_rate, shutter_angle, thread_count, duration)

    def _vend_id(self, alias, otype):  # id_vendor.rs:28-42 (None on collision -> the scene panics)
        if alias in self._aliases:
            raise ValueError(f"This alias collides with another name in the scene! Try changing {alias} to a new name.")
        self._aliases[alias] = (self._next_id, otype)
        self._next_id += 1
        return self._next_id - 1

    def _append(self, kind, data, mat_idx, oid):
        data = np.atleast_2d(np.asarray(data, np.float64))
        n = len(data)
        self._batches.append([kind, data, np.full(n, mat_idx, np.int32), np.full(n, oid, np.int32)])

    def add_element(self, element, alias):  # scene/mod.rs:159-188
        if isinstance(element, Sphere):
            oid = self._vend_id(alias, ObjectType.Sphere)
            self._append(abi.CR_PRIM_SPHERE, np.concatenate([element.center, [element.radius]]),
                         self._tables.material(element.mat), oid)
        elif isinstance(element, Triangle):
            oid = self._vend_id(alias, ObjectType.Triangle)
            self._append(abi.CR_PRIM_TRIANGLE, np.concatenate([element.a, element.b, element.c]),
                         self._tables.material(element.mat), oid)
        elif isinstance(element, Quad):
            oid = self._vend_id(alias, ObjectType.Quad)
            self._append(abi.CR_PRIM_QUAD, np.concatenate([element.q, element.u, element.v]),
                         self._tables.material(element.mat), oid)
        elif isinstance(element, (HitList, BVHWrapper)):
            # scene/mod.rs:160-166: the element is stored as it is: no id is vended, its members keep id 0 (Sphere::new)
            self._append_nested(element)
        else:
            raise TypeError("add_element expects a Sphere, Triangle, Quad, HitList or BVHWrapper")

    def _append_nested(self, element):
        if isinstance(element, BVHWrapper):
            kind, members = abi.CR_GROUP_BVH, element.list.objs
        else:
            kind, members = abi.CR_GROUP_HITLIST, element.objs
        self._batches.append([abi.GROUP_BEGIN, np.zeros((0, 1)), np.array([kind], np.int32), np.zeros(0, np.int32)])
        for m in members:
            if isinstance(m, Sphere):
                self._append(abi.CR_PRIM_SPHERE, np.concatenate([m.center, [m.radius]]), self._tables.material(m.mat), 0)
            elif isinstance(m, Triangle):
                self._append(abi.CR_PRIM_TRIANGLE, np.concatenate([m.a, m.b, m.c]), self._tables.material(m.mat), 0)
            elif isinstance(m, Quad):
                self._append(abi.CR_PRIM_QUAD, np.concatenate([m.q, m.u, m.v]), self._tables.material(m.mat), 0)
            elif isinstance(m, (HitList, BVHWrapper)):
                self._append_nested(m)
            else:
                raise TypeError("a nested list holds Spheres, Triangles, Quads, HitLists or BVHWrappers")
        self._batches.append([abi.GROUP_END, np.zeros((0, 1)), np.zeros(0, np.int32), np.zeros(0, np.int32)])

    def add_spheres(self, centers_radii, mats, alias_prefix):
        """Batch form of add_element for generated scenes: one alias/id per sphere, insertion order kept."""
        cr = np.asarray(centers_radii, np.float64)
        mat_idx = np.array([self._tables.material(m) for m in mats], np.int32)
        oids = np.array([self._vend_id(f"{alias_prefix}{i}", ObjectType.Sphere) for i in range(len(cr))], np.int32)
        self._batches.append([abi.CR_PRIM_SPHERE, cr, mat_idx, oids])

    def load_mesh(self, vertices, faces, alias, scale, shift, mat):
        """Scene::load_asset (scene/mod.rs:191-230) after the OBJ text has been parsed
        (asset_loader/obj_loader.rs:21-53): vertex = scale * p + shift, faces are 1-based, every triangle
        of the mesh shares one id."""
        oid = self._vend_id(alias, ObjectType.TriangleMesh)
        v = scale * np.asarray(vertices, np.float64) + np.asarray(shift, np.float64)
        f = np.asarray(faces, np.int64) - 1
        if f.size and (f.min() < 0 or f.max() >= len(v)):  # obj_loader.rs indexes vertices[i - 1]: out of range panics
            raise IndexError("face index out of range (OBJ face indices are 1-based)")
        tris = np.concatenate([v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]], axis=1)
        self._append(abi.CR_PRIM_TRIANGLE, tris, self._tables.material(mat), oid)

    def load_asset(self, asset_path, alias, scale, shift, mat):
        v, f = parse_obj(asset_path)
        self.load_mesh(v, f, alias, scale, shift, mat)

    def hide_element(self, alias):  # scene/mod.rs:232-281
        self._set_visibility(alias, True)

    def show_element(self, alias):
        self._set_visibility(alias, False)

    def _set_visibility(self, alias, hide):
        if alias not in self._aliases:
            print(f"WARNING: The element `{alias}` does not exist. Are you sure you typed the right name?")
            return
        oid = self._aliases[alias][0]
        (self._hidden_ids.add if hide else self._hidden_ids.discard)(oid)

    def load_default_skybox(self):
        self.sky_kind, self._sky_rgb8 = abi.CR_SKY_DEFAULT, None

    def load_black_skybox(self):  # EXTENSION
        self.sky_kind, self._sky_rgb8 = abi.CR_SKY_BLACK, None

    def load_spherical_skybox(self, rgb8):
        """Scene::load_spherical_skybox with the decoded image (RTWImage forces to_rgb8, img_loader.rs:28)."""
        self.sky_kind, self._sky_rgb8 = abi.CR_SKY_SPHERICAL, np.ascontiguousarray(rgb8, np.uint8)

    # object animation bindings, scene_animator.rs:12-458.  The reference rewrites the timeline of every element
    # carrying the alias' id; here the calls are logged per id and replayed per primitive point in describe().
    def _check_and_get_alias(self, alias, invalid_types, error_msg):  # scene_animator.rs:13-31
        if alias not in self._aliases:
            raise ValueError(f"Could not find an object with the alias: `{alias}`. Are you sure you spelled it right?")
        oid, otype = self._aliases[alias]
        if otype in invalid_types:
            raise ValueError(error_msg)
        return oid

    def scale_r(self, r, keyframe, it, alias):  # scene_animator.rs:140-175
        oid = self._check_and_get_alias(alias, (ObjectType.Camera, ObjectType.Triangle, ObjectType.TriangleMesh, ObjectType.Quad),
                                        "ScaleR can only be applied to Spheres")
        self._anim_ops.setdefault(oid, []).append(("scale_r", float(r), float(keyframe), it))

    def _scale_axis(self, name, kind, v, keyframe, it, alias):  # scene_animator.rs:38-138
        # `invalid_types = [ObjectType::Sphere]`; only Triangle elements carrying the id are rewritten (:50-63)
        oid = self._check_and_get_alias(alias, (ObjectType.Sphere, ObjectType.Quad), f"{name} cannot apply to Spheres")
        self._anim_ops.setdefault(oid, []).append(("scale", kind, float(v), float(keyframe), it))

    def scale_x(self, x, keyframe, it, alias):  # scene_animator.rs:38-67
        self._scale_axis("ScaleX", 4, x, keyframe, it, alias)

    def scale_y(self, y, keyframe, it, alias):  # :72-101
        self._scale_axis("ScaleY", 5, y, keyframe, it, alias)

    def scale_z(self, z, keyframe, it, alias):  # :106-135
        self._scale_axis("ScaleZ", 6, z, keyframe, it, alias)

    def scale_point(self, p, keyframe, it, alias):  # :187-214 -> scale_x, scale_y, scale_z on every vertex timeline
        oid = self._check_and_get_alias(alias, (ObjectType.Sphere, ObjectType.Quad), "ScaleAll cannot apply to Spheres")
        for kind in (4, 5, 6):
            self._anim_ops.setdefault(oid, []).append(("scale", kind, float(p[kind - 4]), float(keyframe), it))

    def scale_all_uniform(self, v, keyframe, it, alias):  # :217-219
        self.scale_point(Point3(v, v, v), keyframe, it, alias)

    def _translate_axis(self, axis, x, keyframe, it, space, alias):
        oid = self._check_and_get_alias(alias, (ObjectType.Quad,), "quads (extension) cannot be animated")
        self._anim_ops.setdefault(oid, []).append(("translate", axis, float(x), float(keyframe), it, space))

    def translate_x(self, x, keyframe, it, space, alias):  # scene_animator.rs:231-282
        self._translate_axis(0, x, keyframe, it, space, alias)

    def translate_y(self, y, keyframe, it, space, alias):  # :284-335
        self._translate_axis(1, y, keyframe, it, space, alias)

    def translate_z(self, z, keyframe, it, space, alias):  # :337-388
        self._translate_axis(2, z, keyframe, it, space, alias)

    def translate_point(self, p, keyframe, it, space, alias):  # :390-458 -> translate_x, _y, _z on every point
        for axis in range(3):
            self._translate_axis(axis, p[axis], keyframe, it, space, alias)

    def _animation(self):
        """[(prim_index, point, [CrAnimKey])] for every animated point, replaying the logged calls on a
        TransformTimeline built from that point's construction position (World-space deltas differ per vertex)."""
        out = []
        if not self._anim_ops:
            return out
        base, depth = 0, 0
        for kind, data, _, oid in self._batches:
            # the scene's id-based calls only see top-level elements (scene_animator.rs:44-59 passes nested ones through)
            depth += 1 if kind == abi.GROUP_BEGIN else (-1 if kind == abi.GROUP_END else 0)
            if kind < 0 or depth > 0:
                base += len(data)
                continue
            for oid_val in np.unique(oid):
                ops = self._anim_ops.get(int(oid_val))
                if not ops:
                    continue
                for row in np.nonzero(oid == oid_val)[0]:
                    pts = [data[row, 0:3]] if kind == abi.CR_PRIM_SPHERE else [data[row, 0:3], data[row, 3:6], data[row, 6:9]]
                    for point, pos in enumerate(pts):
                        tl = TransformTimeline(pos, data[row, 3] if kind == abi.CR_PRIM_SPHERE else 1.0)
                        for op in ops:
                            if op[0] == "scale_r":
                                tl.scale_sphere(op[1], op[2], op[3])
                            elif op[0] == "scale":
                                tl._scale(op[1], op[2], op[3], op[4])
                            else:
                                tl._translate(op[1], op[2], op[3], op[4], op[5])
                        out.append((base + int(row), point, tl.anim_keys()))
            base += len(data)
        return out

    # camera animation bindings, scene_animator.rs:460-552
    def cam_translate_point(self, p, keyframe, interp, space, which):
        tl = self.scene_cam.look_from_tl if which == "from" else self.scene_cam.look_at_tl
        tl.translate_point(p, keyframe, interp, space)

    def cam_translate_x(self, x, keyframe, interp, space, which):
        (self.scene_cam.look_from_tl if which == "from" else self.scene_cam.look_at_tl).translate_x(x, keyframe, interp, space)

    def cam_translate_y(self, y, keyframe, interp, space, which):
        (self.scene_cam.look_from_tl if which == "from" else self.scene_cam.look_at_tl).translate_y(y, keyframe, interp, space)

    def cam_translate_z(self, z, keyframe, interp, space, which):
        (self.scene_cam.look_from_tl if which == "from" else self.scene_cam.look_at_tl).translate_z(z, keyframe, interp, space)

    def compute_frame_count(self):  # scene/mod.rs:324-330
        return int(math.ceil(self.duration * self.frame_rate))

    def describe(self) -> SceneDesc:
        d = SceneDesc()
        d.batches = [tuple(b) for b in self._batches]
        d.materials = list(self._tables.materials)
        d.textures = list(self._tables.textures)
        d.images = list(self._tables.images)
        d.sky_kind = self.sky_kind
        d.animation = self._animation()
        if self._sky_rgb8 is not None:
            d.images.append(self._sky_rgb8)
            d.sky_image = len(d.images) - 1
        if self._hidden_ids:
            base, depth = 0, 0
            for kind, data, _, oid in self._batches:
                depth += 1 if kind == abi.GROUP_BEGIN else (-1 if kind == abi.GROUP_END else 0)
                if kind >= 0 and depth == 0:  # hide_element matches top-level elements only (scene/mod.rs:246-262)
                    hid = np.nonzero(np.isin(oid, list(self._hidden_ids)))[0]
                    d.hidden.extend((base + hid).tolist())
                base += len(data)
        return d

    def render_scene(self, fname, **kw):  # scene/mod.rs:283-347
        from . import gpu

        return gpu.render_scene(self, fname, **kw)


def parse_obj(path):
    """asset_loader/obj_loader.rs:64-143: only `v x y z` and `f a b c` lines; anything else is an error."""
    verts, faces = [], []
    with open(path) as fh:
        for line in fh:
            parts = line.split()
            if not parts:
                continue
            if parts[0] == "v":
                if len(parts) != 4:
                    raise ValueError("Invalid number of coordinates for a vertex")
                verts.append([float(x) for x in parts[1:]])
            elif parts[0] == "f":
                if len(parts) != 4:
                    raise ValueError("The asset loader only supports triangularized images, please triangulate the image then try again")
                faces.append([int(x) for x in parts[1:]])
            else:
                raise ValueError("Unsupported OBJ file")
    return np.array(verts, np.float64), np.array(faces, np.int64)
