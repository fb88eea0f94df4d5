"""N>1 path: row sharding + framebuffer gather.  CPU: world_size-2 gloo run of the gather/assembly host
logic with the oracle standing in for the per-rank render.  GPU: the union of sharded renders is
bit-identical to the single-GPU render (RNG keyed by the global pixel index)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crucible_b200 import demo_builder, multigpu
        from crucible_b200.gpu import rows_of_rank
        from oracle import binding as oracle

        sc = demo_builder.book1_end_scene(image_width=64, samples=2)
        cam = sc.scene_cam.to_abi()
        orc = oracle.OracleScene(sc.describe())
        H, block = cam.image_height, 8
        full_ref, _, _ = orc.render(cam, seed=4, threads=2)
        # this rank's rows only (the oracle stands in for the GPU render of the shard)
        rows = rows_of_rank(H, block, rank, world)
        mine = torch.from_numpy(np.ascontiguousarray(full_ref[rows]))
        out = multigpu.gather_rows(mine, H, block, rank, world)
        frames = multigpu.frames_of_rank(10, rank, world)
        ok = True
        if rank == 0:
            ok = out is not None and np.array_equal(out.numpy(), full_ref)
        else:
            ok = out is None
        q.put((rank, bool(ok), frames))
    finally:
        dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == [0, 2, 4, 6, 8] and res[1][2] == [1, 3, 5, 7, 9]


@pytest.mark.parametrize("H,block,world", [(36, 8, 2), (1080, 8, 8), (37, 8, 3), (5, 8, 4), (64, 1, 3), (2160, 16, 5)])
def test_gather_map_matches_row_sharding(H, block, world):
    """The one-kernel assembly: position of every global row in the stacked [world][max_rows] gather buffer."""
    from crucible_b200 import multigpu
    from crucible_b200.gpu import rows_of_rank

    counts = [len(rows_of_rank(H, block, r, world)) for r in range(world)]
    max_rows = max(counts)
    stacked = np.full((world, max_rows), -1, np.int64)
    for r in range(world):
        stacked[r, : counts[r]] = rows_of_rank(H, block, r, world)  # what rank r sends, padded
    m = multigpu._gather_map(H, block, world, max_rows, "cpu").numpy()
    assert np.array_equal(stacked.reshape(-1)[m], np.arange(H))


def test_gather_rows_world1_is_identity():
    from crucible_b200 import multigpu

    t = torch.arange(24.0).reshape(4, 2, 3)
    assert multigpu.gather_rows(t, 4, 8, 0, 1) is t


@pytest.mark.gpu
def test_sharded_union_equals_single_render(gpu_device):
    """Emulates ranks 0..3 one after the other on one GPU (no kernels wait on each other)."""
    from crucible_b200 import demo_builder
    from crucible_b200.gpu import GpuScene, rows_of_rank

    sc = demo_builder.book1_end_scene(image_width=160, samples=4)
    gs = GpuScene(sc.describe(), gpu_device)
    cam = sc.scene_cam.to_abi()
    full, full8, st = gs.render(cam, seed=6)
    for world in (2, 4):
        acc = np.zeros_like(full)
        acc8 = np.zeros_like(full8)
        rays = 0
        for rank in range(world):
            _, _, s = gs.render(cam, seed=6, row_rank=rank, row_world=world, row_block=8, out_rgb=acc, out_rgb8=acc8)
            rays += s["rays"]
        assert np.array_equal(acc, full) and np.array_equal(acc8, full8)
        assert rays == st["rays"]


@pytest.mark.gpu
def test_render_device_packed_rows(gpu_device):
    from crucible_b200 import demo_builder, multigpu
    from crucible_b200.gpu import GpuScene, rows_of_rank

    sc = demo_builder.book1_end_scene(image_width=96, samples=2)
    gs = GpuScene(sc.describe(), gpu_device)
    cam = sc.scene_cam.to_abi()
    full, full8, _ = gs.render(cam, seed=8)
    # world 1 through the torch path
    t, t8, _ = multigpu.render_sharded(gs, cam, 0, 1, seed=8)
    assert np.array_equal(t.cpu().numpy(), full) and np.array_equal(t8.cpu().numpy(), full8)
    # one shard of three, packed rows straight into a torch tensor
    rows = rows_of_rank(cam.image_height, 8, 1, 3)
    buf = torch.zeros((len(rows), cam.image_width, 3), dtype=torch.float64, device="cuda")
    gs.render_device(cam, buf.data_ptr(), 0, stream=torch.cuda.current_stream().cuda_stream, seed=8, row_rank=1, row_world=3)
    assert np.array_equal(buf.cpu().numpy(), full[rows])


@pytest.mark.gpu
def test_animation_frames_sharded(gpu_device, oracle):
    """BASELINE config 5 in miniature: whole frames sharded over ranks (frame f -> rank f % world), camera
    keyframes evaluated on the device at every sample time; each frame equals the oracle's."""
    from crucible_b200 import demo_builder, multigpu
    from crucible_b200.gpu import GpuScene

    sc = demo_builder.book1_walkthrough(image_width=64, samples=2, duration=0.5)  # 12 frames
    assert sc.compute_frame_count() == 12
    desc = sc.describe()
    gs, orc = GpuScene(desc, gpu_device), oracle.OracleScene(desc)
    got = {}
    for rank in range(3):  # three ranks emulated one after the other
        multigpu.render_frames_sharded(gs, sc, rank, 3, seed=4, on_frame=lambda f, img: got.__setitem__(f, img.copy()))
    assert sorted(got) == list(range(12))
    cam = sc.scene_cam.to_abi()
    for f in (0, 5, 11):
        cam.frame = f
        ref, ref8, _ = orc.render(cam, seed=4)
        assert (got[f] != ref8).mean() < 0.01  # bytes differ only at razor-edge values (see test_render_parity)
    assert not np.array_equal(got[0], got[11])  # the camera really moved


@pytest.mark.gpu
def test_render_multi_equals_single_render(gpu_device, crlib):
    """cr_render_multi over every visible device (one replica per device, rows stored into the first device's image by
    each device's resolve kernel): bit-identical to cr_render on one device.  With one device this still exercises
    cr_scene_replicate and the global-row device output."""
    from crucible_b200 import abi, demo_builder, gpu
    from crucible_b200.gpu import GpuScene, rows_of_rank

    sc = demo_builder.load_teapot(image_width=128, samples=3)
    gs = GpuScene(sc.describe(), gpu_device)
    cam = sc.scene_cam.to_abi()
    full, full8, st = gs.render(cam, seed=12)
    n = min(crlib.cr_device_count(), 8)
    replicas = [gs] + [gs.replicate(d) for d in range(1, n)]
    assert replicas[-1].bvh_info() == gs.bvh_info()
    rgb, rgb8, stats = gpu.render_multi(replicas, cam, seed=12)
    assert np.array_equal(rgb, full) and np.array_equal(rgb8, full8)
    assert sum(s["rays"] for s in stats) == st["rays"] and len(stats) == n
    _, only8, _ = gpu.render_multi(replicas, cam, seed=12, want_rgb=False)  # bytes only: what Camera::render consumes
    assert np.array_equal(only8, full8)
    # a replica on the same device is a second, independent scene handle
    twin = gs.replicate(gpu_device)
    a, _, _ = twin.render(cam, seed=12)
    assert np.array_equal(a, full)
    with pytest.raises(abi.CrucibleError, match="two replicas on one device"):
        gpu.render_multi([gs, twin], cam, seed=12)
    # global-row device output of one shard: the other rows stay untouched
    H, W = cam.image_height, cam.image_width
    buf = torch.full((H, W, 3), -1.0, dtype=torch.float64, device="cuda")
    gs.render_device(cam, buf.data_ptr(), 0, stream=torch.cuda.current_stream().cuda_stream, seed=12, row_rank=1, row_world=3,
                     global_rows=True)
    rows = rows_of_rank(H, 8, 1, 3)
    out = buf.cpu().numpy()
    assert np.array_equal(out[rows], full[rows])
    rest = np.setdiff1d(np.arange(H), rows)
    assert (out[rest] == -1.0).all()
    for r in replicas[1:] + [twin]:
        r.close()
