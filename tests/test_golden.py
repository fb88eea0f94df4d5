"""The oracle against its own frozen outputs (tests/golden/, scripts/make_golden.py): closest-hit records of seeded
SURVEY 8d ray batches and one small render per config.  Bit for bit: same code, same libm, same machine image."""
import numpy as np
import pytest

import golden


@pytest.mark.parametrize("name", list(golden.CONFIGS))
def test_oracle_reproduces_golden(oracle, name):
    z = golden.load(name)
    sc = golden.build(name)
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    orc = oracle.OracleScene(desc)
    info = orc.bvh_info()
    assert [info[k] for k in ("n_nodes", "max_depth", "n_visible")] == z["bvh"].tolist()
    for bname, rays in golden.batches(desc, cam, orc).items():
        idx = z[bname + "_idx"]
        assert np.array_equal(idx, golden.keep_indices(len(rays)))
        assert np.array_equal(rays[idx], z[bname + "_rays"]), bname  # camera rays / first-bounce rays / random rays
        got, exp = orc.trace_batch(z[bname + "_rays"]), z[bname + "_hits"]
        assert got.tobytes() == exp.tobytes(), bname
        assert (exp["prim_index"] >= 0).sum() > 50, bname
    rgb, rgb8, st = orc.render(golden.small_camera(name), seed=golden.RENDER["seed"])
    assert np.array_equal(rgb, z["render_rgb"]) and np.array_equal(rgb8, z["render_rgb8"])
    assert st["rays"] == int(z["render_rays"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(golden.CONFIGS))
def test_cuda_path_reproduces_golden(gpu_device, name):
    """The CUDA path against the FROZEN records alone (no live oracle in this test): ids / t / p / n bit-exact,
    the small render within 1e-11 and the same number of world.hit calls."""
    from crucible_b200.gpu import GpuScene

    z = golden.load(name)
    sc = golden.build(name)
    gs = GpuScene(sc.describe(), gpu_device)
    info = gs.bvh_info()
    assert [info[k] for k in ("n_nodes", "max_depth", "n_visible")] == z["bvh"].tolist()
    for bname in ("primary", "first_bounce", "random"):
        rays, exp = z[bname + "_rays"], z[bname + "_hits"]
        got = gs.trace_batch(rays)
        for f in ("prim_index", "obj_id", "front_face", "material"):
            assert np.array_equal(got[f], exp[f]), (bname, f)
        hit = exp["prim_index"] >= 0
        for f in ("t", "p", "n"):
            assert np.array_equal(got[f][hit], exp[f][hit]), (bname, f)
        assert np.all(np.abs(got["u"][hit] - exp["u"][hit]) <= 1e-5) and np.all(np.abs(got["v"][hit] - exp["v"][hit]) <= 1e-5)
    rgb, _, st = gs.render(golden.small_camera(name), seed=golden.RENDER["seed"])
    # image-textured spheres (earthmap on config 4's spheres) may flip a texel where u*W is within an ulp of an integer
    diff = np.abs(rgb - z["render_rgb"])
    assert (diff.max(axis=2) > 1e-11).mean() <= (1e-3 if name == "instanced64" else 0.0)
    assert st["rays"] == int(z["render_rays"][0])
