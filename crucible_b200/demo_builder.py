"""Seeded restatements of the reference's demo scenes (src/demo_builder/demo_images.rs, demo_movies.rs)
and of the five BASELINE.json configurations.

The reference draws its scene from an UNSEEDED thread RNG (demo_images.rs:45), so every reference run
renders a different book1.  Here the same draws are made in the same order from a seeded Philox
generator, which makes the scene a reproducible test/bench input.
"""
from __future__ import annotations

import os

import numpy as np

from . import abi
from .scene import (CheckerTexture, Color, Dielectric, Emissive, ImageTexture, InterpolationType, Lambertian, Metal, Point3,
                    Quad, Scene, Sphere, TransformSpace)

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=int(seed)))


def teapot_mesh():
    z = np.load(os.path.join(_ASSETS, "teapot.npz"))
    return z["vertices"], z["faces"].astype(np.int64)


def earthmap_rgb8():
    return np.load(os.path.join(_ASSETS, "earthmap.npz"))["rgb8"]


def garden_substitute_rgb8(w=4096, h=2048):
    """assets/garden.hdr is missing from the reference checkout (.MISSING_LARGE_BLOBS).  Substitute: a fixed
    procedural equirect (sky gradient + sun lobe + ground), clamped and quantised to RGB8 exactly as the
    reference's to_rgb8() would do to a decoded HDR (img_loader.rs:28)."""
    v = (np.arange(h) + 0.5) / h  # 0 = top row
    u = (np.arange(w) + 0.5) / w
    phi = (0.5 - v) * np.pi  # elevation
    theta = (u - 0.5) * 2 * np.pi
    el = phi[:, None]
    az = theta[None, :]
    sky = np.stack([0.35 + 0.4 * (1 - np.sin(np.clip(el, 0, None))) + 0 * az,
                    0.55 + 0.3 * (1 - np.sin(np.clip(el, 0, None))) + 0 * az,
                    0.95 + 0 * el + 0 * az], axis=-1)
    ground = np.stack([0.25 + 0.05 * np.cos(7 * az) + 0 * el, 0.32 + 0.05 * np.sin(5 * az) + 0 * el, 0.12 + 0 * el + 0 * az], axis=-1)
    img = np.where((el >= 0)[..., None], sky, ground)
    sun_dir = np.array([np.cos(0.7) * np.sin(1.0), np.sin(0.7), np.cos(0.7) * np.cos(1.0)])
    d = np.stack([np.cos(el) * np.sin(az), np.sin(el) + 0 * az, np.cos(el) * np.cos(az)], axis=-1)
    lobe = np.clip((d @ sun_dir), 0, 1) ** 256
    img = img + 4.0 * lobe[..., None]
    return (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def _checker_ground():
    checker = CheckerTexture.new_from_color(0.32, Color(0.2, 0.3, 0.1), Color(0.9, 0.9, 0.9))
    return Lambertian.new_from_texture(checker, 1.0)


def book1_end_scene(threads=0, seed=1, image_width=400, samples=500):
    """demo_images.rs:14-109.  BASELINE config 1 = image_width 1920, samples 100."""
    sc = Scene.new_image(16.0 / 9.0, image_width, 24, 180.0, threads)
    cam = sc.scene_cam
    cam.set_samples(samples)
    cam.set_max_depth(50)
    cam.look_from(Point3(13.0, 2.0, 3.0))
    cam.look_at(Point3(0.0, 0.0, 0.0))
    cam.set_vfov(20.0)
    cam.set_defocus_angle(0.6)
    cam.set_focus_dist(10.0)
    sc.add_element(Sphere(Point3(0.0, -1000.0, 0.0), 1000.0, _checker_ground()), "ground")
    rng = _rng(seed)
    centers, mats = [], []
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose_mat = rng.random()
            center = Point3(a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
            d = center - Point3(4.0, 0.2, 0.0)
            if np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.9:
                if choose_mat < 0.8:
                    c1 = Color(rng.random(), rng.random(), rng.random())
                    c2 = Color(rng.random(), rng.random(), rng.random())
                    mat = Lambertian.new_from_color(c1 * c2, 1.0)
                elif choose_mat < 0.95:
                    albedo = Color(*[0.5 + 0.5 * rng.random() for _ in range(3)])
                    mat = Metal(albedo, 0.5 * rng.random())
                else:
                    mat = Dielectric(1.5)
                centers.append([center[0], center[1], center[2], 0.2])
                mats.append(mat)
    sc.add_spheres(np.array(centers), mats, "small")
    sc.add_element(Sphere(Point3(0.0, 1.0, 0.0), 1.0, Dielectric(1.5)), "large_dielectric")
    sc.add_element(Sphere(Point3(-4.0, 1.0, 0.0), 1.0, Lambertian.new_from_color(Color(0.4, 0.2, 0.1), 1.0)), "large_lambertian")
    sc.add_element(Sphere(Point3(4.0, 1.0, 0.0), 1.0, Metal(Color(0.7, 0.6, 0.5), 0.0)), "large_metal")
    return sc


def checkered_spheres(threads=0, image_width=400, samples=500):
    """demo_images.rs:112-152."""
    sc = Scene.new_image(16.0 / 9.0, image_width, 24, 180.0, threads)
    cam = sc.scene_cam
    cam.set_samples(samples)
    cam.set_max_depth(50)
    cam.look_from(Point3(13.0, 2.0, 3.0))
    cam.look_at(Point3(0.0, 0.0, 0.0))
    cam.set_vfov(20.0)
    cam.set_defocus_angle(0.6)
    cam.set_focus_dist(10.0)
    checker = CheckerTexture.new_from_color(0.32, Color(0.2, 0.3, 0.1), Color(0.9, 0.9, 0.9))
    sc.add_element(Sphere(Point3(0.0, -10.0, 0.0), 10.0, Lambertian.new_from_texture(checker, 1.0)), "bottom_sphere")
    sc.add_element(Sphere(Point3(0.0, 10.0, 0.0), 10.0, Lambertian.new_from_texture(checker, 1.0)), "top_sphere")
    return sc


def load_teapot(threads=0, image_width=400, samples=200, sky="garden"):
    """demo_images.rs:155-200 (+ spherical sky for BASELINE config 3: image_width 1920, samples 256)."""
    sc = Scene.new_image(16.0 / 9.0, image_width, 24, 180.0, threads)
    cam = sc.scene_cam
    cam.set_samples(samples)
    cam.set_max_depth(50)
    cam.look_from(Point3(13.0, 10.0, 3.0))
    cam.look_at(Point3(0.0, 0.0, 0.0))
    cam.set_vfov(20.0)
    cam.set_defocus_angle(0.6)
    cam.set_focus_dist(10.0)
    v, f = teapot_mesh()
    sc.load_mesh(v, f, "teapot", 0.5, Point3(0.0, 0.0, 0.0), Metal(Color(0.8, 0.3, 0.5), 0.05))
    sc.add_element(Sphere(Point3(0.0, -1000.0, 0.0), 1000.0, _checker_ground()), "ground")
    if sky == "garden":
        sc.load_spherical_skybox(garden_substitute_rgb8())
    return sc


def earth(threads=0, image_width=400, samples=500):
    """demo_images.rs:202-221."""
    sc = Scene.new_image(16.0 / 9.0, image_width, 24, 180.0, threads)
    cam = sc.scene_cam
    cam.set_samples(samples)
    cam.set_max_depth(50)
    cam.look_from(Point3(0.0, 0.0, 12.0))
    cam.look_at(Point3(0.0, 0.0, 0.0))
    cam.set_vfov(20.0)
    tex = ImageTexture(earthmap_rgb8())
    sc.add_element(Sphere(Point3(0.0, 0.0, 0.0), 2.0, Lambertian.new_from_texture(tex, 1.0)), "earth")
    return sc


def garden_skybox(threads=0, image_width=1920, samples=500):
    """demo_images.rs:223-242 with the procedural garden substitute."""
    sc = Scene.new_image(16.0 / 9.0, image_width, 24, 180.0, threads)
    cam = sc.scene_cam
    cam.set_samples(samples)
    cam.set_max_depth(50)
    cam.look_from(Point3(0.0, 0.0, -12.0))
    cam.look_at(Point3(0.0, 0.0, 0.0))
    cam.set_vfov(40.0)
    sc.add_element(Sphere(Point3(0.0, 0.0, 0.0), 2.0, Metal(Color(0.8, 0.8, 0.8), 0.05)), "metal_ball")
    sc.load_spherical_skybox(garden_substitute_rgb8())
    return sc


def _box_quads(a, b, mat):
    """RTNW `box(a, b)`: six quads of the axis-aligned box with opposite corners a, b."""
    mn, mx = np.minimum(a, b), np.maximum(a, b)
    dx, dy, dz = Point3(mx[0] - mn[0], 0, 0), Point3(0, mx[1] - mn[1], 0), Point3(0, 0, mx[2] - mn[2])
    return [Quad(Point3(mn[0], mn[1], mx[2]), dx, dy, mat), Quad(Point3(mx[0], mn[1], mx[2]), -dz, dy, mat),
            Quad(Point3(mx[0], mn[1], mn[2]), -dx, dy, mat), Quad(Point3(mn[0], mn[1], mn[2]), dz, dy, mat),
            Quad(Point3(mn[0], mx[1], mx[2]), dx, -dz, mat), Quad(Point3(mn[0], mn[1], mn[2]), dx, dz, mat)]


def cornell_box(threads=0, image_width=1024, samples=1000):
    """BASELINE config 2 (EXTENSION: Quad + Emissive do not exist in the reference): RTNW Cornell box,
    555^3, light quad (343,554,332)+(-130,0,0),(0,0,-105) emitting 15, two UNROTATED boxes, black sky."""
    sc = Scene.new_image(1.0, image_width, 24, 180.0, threads)
    cam = sc.scene_cam
    cam.set_samples(samples)
    cam.set_max_depth(50)
    cam.look_from(Point3(278.0, 278.0, -800.0))
    cam.look_at(Point3(278.0, 278.0, 0.0))
    cam.set_vfov(40.0)
    red = Lambertian.new_from_color(Color(0.65, 0.05, 0.05), 1.0)
    white = Lambertian.new_from_color(Color(0.73, 0.73, 0.73), 1.0)
    green = Lambertian.new_from_color(Color(0.12, 0.45, 0.15), 1.0)
    light = Emissive((15.0, 15.0, 15.0))
    quads = [("green_wall", Quad(Point3(555, 0, 0), Point3(0, 555, 0), Point3(0, 0, 555), green)),
             ("red_wall", Quad(Point3(0, 0, 0), Point3(0, 555, 0), Point3(0, 0, 555), red)),
             ("light", Quad(Point3(343, 554, 332), Point3(-130, 0, 0), Point3(0, 0, -105), light)),
             ("floor", Quad(Point3(0, 0, 0), Point3(555, 0, 0), Point3(0, 0, 555), white)),
             ("ceiling", Quad(Point3(555, 555, 555), Point3(-555, 0, 0), Point3(0, 0, -555), white)),
             ("back", Quad(Point3(0, 0, 555), Point3(555, 0, 0), Point3(0, 555, 0), white))]
    for alias, q in quads:
        sc.add_element(q, alias)
    for k, q in enumerate(_box_quads(Point3(130, 0, 65), Point3(295, 165, 230), white)):
        sc.add_element(q, f"box1_{k}")
    for k, q in enumerate(_box_quads(Point3(265, 0, 295), Point3(430, 330, 460), white)):
        sc.add_element(q, f"box2_{k}")
    sc.load_black_skybox()
    return sc


def instanced_teapots(threads=0, image_width=3840, samples=64, copies=1582, grid=40, spacing=4.0):
    """BASELINE config 4: `copies` x load_asset("teapot.obj") on a grid (flattened exactly as
    scene/mod.rs:211-229 does: no instancing in the reference), materials cycling
    Lambertian(earthmap) / metal / glass, + 64 earth-textured r=1 spheres + checker ground.
    copies=1582 -> 9 998 240 triangles."""
    sc = Scene.new_image(16.0 / 9.0, image_width, 24, 180.0, threads)
    cam = sc.scene_cam
    cam.set_samples(samples)
    cam.set_max_depth(50)
    cam.look_from(Point3(60.0, 45.0, 60.0))
    cam.look_at(Point3(0.0, 0.0, 0.0))
    cam.set_vfov(40.0)
    v, f = teapot_mesh()
    earth_tex = ImageTexture(earthmap_rgb8())
    mats = [Lambertian.new_from_texture(earth_tex, 1.0), Metal(Color(0.8, 0.6, 0.3), 0.05), Dielectric(1.5)]
    half = (grid - 1) * spacing / 2.0
    for k in range(copies):
        gx, gz = k % grid, k // grid
        shift = Point3(gx * spacing - half, 0.0, gz * spacing - half)
        sc.load_mesh(v, f, f"teapot{k}", 0.5, shift, mats[k % 3])
    earth_mat = Lambertian.new_from_texture(earth_tex, 1.0)
    for k in range(64):
        gx, gz = k % 8, k // 8
        sc.add_element(Sphere(Point3(gx * 20.0 - 70.0, 3.0, gz * 20.0 - 70.0), 1.0, earth_mat), f"earth{k}")
    sc.add_element(Sphere(Point3(0.0, -1000.0, 0.0), 1000.0, _checker_ground()), "ground")
    return sc


def book1_walkthrough(threads=0, seed=1, image_width=1920, samples=64, frame_rate=24, duration=10.0):
    """BASELINE config 5: the book1 scene with camera keyframes in the pattern of demo_movies.rs:33-68."""
    sc = book1_end_scene(threads, seed, image_width, samples)
    sc.duration = duration
    for p, kf in ((Point3(3.0, 2.0, 13.0), 2.5), (Point3(-13.0, 2.0, 3.0), 5.0), (Point3(-3.0, 2.0, -13.0), 7.5),
                  (Point3(13.0, 2.0, 3.0), 10.0)):
        sc.cam_translate_point(p, kf, InterpolationType.LERP, TransformSpace.World, "from")
    return sc


def first_movie(threads=0, frame_rate=24, duration=10.0, image_width=400):
    """demo_movies.rs:12-70 with the procedural garden substitute."""
    sc = Scene.new_movie(16.0 / 9.0, image_width, frame_rate, 180.0, threads, duration)
    cam = sc.scene_cam
    cam.set_samples(50)
    cam.set_max_depth(5)
    cam.look_from(Point3(0.0, 0.0, -12.0))
    cam.look_at(Point3(0.0, 0.0, 0.0))
    cam.set_vfov(40.0)
    sc.add_element(Sphere(Point3(0.0, 0.0, 0.0), 2.0, Metal(Color(0.8, 0.8, 0.8), 0.05)), "metal_ball")
    sc.load_spherical_skybox(garden_substitute_rgb8(1024, 512))
    for p, kf in ((Point3(12.0, 0.0, 0.0), 2.5), (Point3(0.0, 0.0, 12.0), 5.0), (Point3(-12.0, 0.0, 0.0), 7.5),
                  (Point3(0.0, 0.0, -12.0), 10.0), (Point3(0.0, 5.0, -20.0), 15.0)):
        sc.cam_translate_point(p, kf, InterpolationType.LERP, TransformSpace.World, "from")
    return sc


CONFIGS = {
    "book1": lambda **kw: book1_end_scene(image_width=kw.get("image_width", 1920), samples=kw.get("samples", 100), seed=kw.get("seed", 1)),
    "cornell": lambda **kw: cornell_box(image_width=kw.get("image_width", 1024), samples=kw.get("samples", 1000)),
    "teapot": lambda **kw: load_teapot(image_width=kw.get("image_width", 1920), samples=kw.get("samples", 256)),
    "instanced": lambda **kw: instanced_teapots(image_width=kw.get("image_width", 3840), samples=kw.get("samples", 64),
                                                copies=kw.get("copies", 1582), grid=kw.get("grid", 40),
                                                spacing=kw.get("spacing", 4.0)),
    "walkthrough": lambda **kw: book1_walkthrough(image_width=kw.get("image_width", 1920), samples=kw.get("samples", 64),
                                                  seed=kw.get("seed", 1)),
}
