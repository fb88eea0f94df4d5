"""The file-writing tail of the path: Camera::render's PPM (camera/mod.rs:275-311, Display for Color
utils.rs:422-438) and Scene::render_movie's frame loop (scene/mod.rs:295-330)."""
import ctypes as C
import os

import numpy as np
import pytest

from crucible_b200 import abi, demo_builder, gpu


def _read_p3(path):
    tok = open(path).read().split()
    assert tok[0] == "P3" and tok[3] == "255"
    w, h = int(tok[1]), int(tok[2])
    return np.array(tok[4:], dtype=np.int64).reshape(h, w, 3).astype(np.uint8)


def test_write_ppm_p3_is_the_reference_text(tmp_path):
    """Byte for byte the text `writeln!(bw, "P3\\n{iw} {ih}\\n255")` + one `{r} {g} {b}` line per pixel."""
    rng = np.random.default_rng(3)
    for h, w in [(1, 1), (3, 5), (37, 41), (270, 480)]:  # the last one takes the multi-threaded formatter
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        img[0, 0] = (0, 9, 10)
        img[-1, -1] = (99, 100, 255)
        a, b = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm")
        gpu.write_ppm(a, img)
        gpu.write_ppm_python(b, img)
        assert open(a, "rb").read() == open(b, "rb").read()
        assert np.array_equal(_read_p3(a), img)


def test_write_ppm_p6_and_truncation(tmp_path):
    img = np.arange(4 * 6 * 3, dtype=np.uint8).reshape(4, 6, 3)
    path = str(tmp_path / "x.ppm")
    open(path, "wb").write(b"x" * 10000)  # OpenOptions::truncate(true), camera/mod.rs:275-280
    gpu.write_ppm(path, img, abi.CR_PPM_P6)
    raw = open(path, "rb").read()
    assert raw == b"P6\n6 4\n255\n" + img.tobytes()


def test_write_ppm_errors(tmp_path):
    lib = abi.load()
    img = np.zeros((2, 2, 3), np.uint8)
    assert lib.cr_write_ppm(os.fsencode(str(tmp_path / "nodir" / "x.ppm")), img.ctypes.data_as(C.c_void_p), 2, 2, abi.CR_PPM_P3) == abi.CR_ERR_INVALID
    assert b"cannot open" in lib.cr_last_error()
    assert lib.cr_write_ppm(os.fsencode(str(tmp_path / "x.ppm")), img.ctypes.data_as(C.c_void_p), 2, 2, 7) == abi.CR_ERR_INVALID
    assert lib.cr_write_ppm(None, img.ctypes.data_as(C.c_void_p), 2, 2, abi.CR_PPM_P3) == abi.CR_ERR_INVALID


def test_file_entry_points_need_a_device(tmp_path):
    """No CPU fallback: without an sm_100 device the render-to-file calls fail with CR_ERR_NO_DEVICE."""
    lib = abi.load()
    if lib.cr_device_count() > 0:
        pytest.skip("a GPU is present")
    h = lib.cr_scene_create(-1)
    cam = demo_builder.book1_end_scene(image_width=16, samples=1, seed=1).scene_cam.to_abi()
    opts = abi.CrRenderOpts(1, abi.CR_PRECISION_F64, 0, 8, 0, 1, 0)
    assert lib.cr_render_to_file(h, C.byref(cam), C.byref(opts), os.fsencode(str(tmp_path / "a.ppm")), abi.CR_PPM_P3, None) in (
        abi.CR_ERR_NO_DEVICE, abi.CR_ERR_STATE)
    assert lib.cr_render_frames(h, C.byref(cam), C.byref(opts), 0, 1, 2, os.fsencode(str(tmp_path)), 1, abi.CR_PPM_P3, None) in (
        abi.CR_ERR_NO_DEVICE, abi.CR_ERR_STATE)
    lib.cr_scene_destroy(h)


@pytest.mark.gpu
def test_render_scene_writes_the_reference_file(gpu_device, oracle, tmp_path):
    """Scene::render_scene for a still: `<fname>.ppm` holds the bytes of the oracle's render (the f64 path
    follows the oracle's paths; bytes may differ only where 255*sqrt(c) sits on an integer)."""
    sc = demo_builder.book1_end_scene(image_width=96, samples=4, seed=2)
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    st = sc.render_scene(str(tmp_path / "still"), seed=5)
    got = _read_p3(str(tmp_path / "still.ppm"))
    _, ref8, ost = oracle.OracleScene(desc).render(cam, seed=5)
    assert got.shape == ref8.shape
    assert (got != ref8).mean() < 0.01
    assert st[0]["rays"] == ost["rays"]


@pytest.mark.gpu
def test_render_movie_frame_loop(gpu_device, tmp_path):
    """Scene::render_movie: artifacts/imageNN.ppm for every frame, each identical to that frame rendered
    on its own; frame sharding (first/stride) writes disjoint files whose union is the same set; P6 holds
    the same bytes."""
    sc = demo_builder.book1_walkthrough(image_width=64, samples=2, duration=0.5)  # 12 frames at 24 fps
    frames = sc.compute_frame_count()
    assert frames == 12
    desc = sc.describe()
    gs = gpu.GpuScene(desc, gpu_device)
    cam = sc.scene_cam.to_abi()
    want = []
    for f in range(frames):
        cam.frame = f
        _, rgb8, _ = gs.render(cam, seed=4, want_rgb=False)
        want.append(rgb8.copy())
    assert any(not np.array_equal(want[0], w) for w in want[1:])  # the camera moves
    stats = gpu.render_scene(sc, str(tmp_path / "mv"), seed=4, gpu_scene=gs, make_movie=False)
    assert len(stats) == frames and sc.scene_cam.frame == frames
    names = sorted(os.listdir(tmp_path / "mv" / "artifacts"))
    assert names == [f"image{f:02d}.ppm" for f in range(frames)]
    for f in range(frames):
        assert np.array_equal(_read_p3(str(tmp_path / "mv" / "artifacts" / names[f])), want[f]), f
    # sharded over 3 ranks, binary output
    sc2 = demo_builder.book1_walkthrough(image_width=64, samples=2, duration=0.5)
    for rank in range(3):
        out = gpu.render_scene(sc2, str(tmp_path / "sh"), seed=4, gpu_scene=gs, fmt=abi.CR_PPM_P6, rank=rank, world=3, make_movie=False) \
            if rank == 0 else _shard(gs, sc2, str(tmp_path / "sh"), rank, 3)
        assert len(out) == len(range(rank, frames, 3))
    for f in range(frames):
        raw = open(tmp_path / "sh" / "artifacts" / f"image{f:02d}.ppm", "rb").read()
        assert raw == b"P6\n64 36\n255\n" + want[f].tobytes(), f
    gs.close()


def _shard(gs, sc, fname, rank, world):
    """Ranks > 0 of a frame-sharded movie (rank 0 created the directory)."""
    cam = sc.scene_cam.to_abi()
    cam.frame = 0
    opts = abi.CrRenderOpts(4, abi.CR_PRECISION_F64, 0, 8, 0, 1, 0)
    frames = sc.compute_frame_count()
    n = len(range(rank, frames, world))
    stats = (abi.CrStats * n)()
    abi.check(gs.lib.cr_render_frames(gs.handle, C.byref(cam), C.byref(opts), rank, world, frames,
                                      os.fsencode(os.path.join(fname, "artifacts")), len(str(frames)), abi.CR_PPM_P6,
                                      C.cast(stats, C.c_void_p)))
    return [stats[i].as_dict() for i in range(n)]


def test_write_png_round_trip(tmp_path):
    """8-bit RGB PNG with the bytes of the PPM (one zlib stream assembled from independently deflated row
    bands, CRC-32 per chunk); read back with an independent decoder."""
    from PIL import Image

    rng = np.random.default_rng(8)
    for h, w in [(1, 1), (7, 5), (300, 400), (1080, 1920)]:  # the larger ones take several bands
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        img[: h // 2] = (np.arange(w)[None, :, None] % 256).astype(np.uint8)  # a compressible half
        path = str(tmp_path / f"x{h}.png")
        gpu.write_ppm(path, img, abi.CR_PNG)
        raw = open(path, "rb").read()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n" and raw[12:16] == b"IHDR" and raw[-8:-4] == b"IEND"
        back = np.array(Image.open(path).convert("RGB"))
        assert np.array_equal(back, img)
        if h >= 300:
            assert len(raw) < img.nbytes * 0.8  # the compressible half really is compressed


def test_scene_file_round_trip_host(tmp_path, crlib):
    """cr_scene_save / cr_scene_load: primitives in insertion order with materials, ids and hidden flags, tables, images,
    sky, keyframes: the loaded scene builds the same tree (checked without a GPU on a host-only scene)."""
    from crucible_b200.gpu import GpuScene
    from crucible_b200.scene import InterpolationType, TransformSpace

    sc = demo_builder.load_teapot(image_width=64, samples=1)
    sc.hide_element("ground")
    sc.translate_x(0.5, 1.0, InterpolationType.LERP, TransformSpace.Local, "teapot")
    gs = GpuScene(sc.describe(), device=-1)
    path = str(tmp_path / "scene.crs")
    gs.save(path)
    back = GpuScene.load(path, device=-1)
    assert back.bvh_info() == gs.bvh_info() and back.bvh_info()["n_visible"] == 6320
    assert np.array_equal(back.bvh_leaf_order(), gs.bvh_leaf_order())
    a, b = gs.bvh_nodes(), back.bvh_nodes()
    assert a.tobytes() == b.tobytes()
    again = str(tmp_path / "again.crs")
    back.save(again)
    assert open(path, "rb").read() == open(again, "rb").read()  # nothing lost, nothing reordered
    # a damaged file is refused
    raw = bytearray(open(path, "rb").read())
    raw[:4] = b"XXXX"
    open(again, "wb").write(raw)
    with pytest.raises(abi.CrucibleError, match="not a valid scene file"):
        GpuScene.load(again, device=-1)
    open(again, "wb").write(bytes(open(path, "rb").read()[:1000]))  # truncated
    with pytest.raises(abi.CrucibleError, match="not a valid scene file"):
        GpuScene.load(again, device=-1)


@pytest.mark.gpu
def test_scene_file_and_png_on_the_gpu(tmp_path, gpu_device):
    """A scene saved, loaded and rendered gives the image of the original bit for bit; cr_render_to_file writes it as PNG."""
    from PIL import Image

    from crucible_b200.gpu import GpuScene

    sc = demo_builder.load_teapot(image_width=96, samples=3)
    cam = sc.scene_cam.to_abi()
    gs = GpuScene(sc.describe(), gpu_device)
    rgb, rgb8, st = gs.render(cam, seed=2)
    path = str(tmp_path / "scene.crs")
    gs.save(path)
    back = GpuScene.load(path, gpu_device)
    rgb_b, rgb8_b, st_b = back.render(cam, seed=2)
    assert np.array_equal(rgb_b, rgb) and np.array_equal(rgb8_b, rgb8) and st_b["rays"] == st["rays"]
    opts = abi.CrRenderOpts(2, abi.CR_PRECISION_F64, 0, 8, 0, 1, 0)
    png = str(tmp_path / "out.png")
    abi.check(back.lib.cr_render_to_file(back.handle, C.byref(cam), C.byref(opts), os.fsencode(png), abi.CR_PNG, None))
    assert np.array_equal(np.array(Image.open(png).convert("RGB")), rgb8)
