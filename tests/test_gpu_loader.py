"""the loader's train-split sample pipeline on the GPU (harness/gpu_loader.py, csrc/augment.cu)
against the numpy restatement oracle/dataset_ref.py (itself pinned bit-exact against the reference's own
transform.elastic / jitter / flip / rotz in tests/test_cpu_oracle_and_host.py) ON SHARED RANDOM DRAWS.

Tolerances: augmented coordinates 1e-6 (float32 dot product: numpy's sgemm uses FMA in an unspecified order, the
kernel separate multiplies and adds -- one float32 ulp).  That ulp is the input of everything downstream, so the
elastic coordinates (double precision on both sides; the elastic kernels alone match scipy to 1e-9, second test) agree
to 2e-6 m, and the voxel coordinates are identical for all but at most 5e-4 of the points (a point within ~1e-6 m of
a voxel face rounds to the other side); where they are identical the maps and features must be identical too."""
import numpy as np
import pytest
import torch

from oracle import dataset_ref

pytestmark = pytest.mark.gpu


def _scene(seed, n):
    from minsu3d_b200.harness import scenes
    return scenes.make_scene(seed, n)


def _to_dev(sc):
    return {"xyz": torch.from_numpy(sc["xyz"]).cuda(), "rgb": torch.from_numpy(sc["rgb"]).cuda(),
            "sem_labels": torch.from_numpy(sc["sem_labels"]).cuda(), "instance_ids": torch.from_numpy(sc["instance_ids"]).cuda()}


@pytest.mark.parametrize("n,max_pts", [(30_000, 250_000), (100_000, 250_000), (60_000, 20_000)])
def test_gpu_train_sample_matches_numpy_restatement_on_shared_draws(n, max_pts):
    from minsu3d_b200.harness import gpu_loader
    sc = _scene(3, n)
    np.random.seed(100 + n)
    want, draws = dataset_ref.train_sample(sc, max_num_point=max_pts, full_scale=(128, 512))
    got = gpu_loader.train_sample_gpu(_to_dev(sc), {k: (list(v) if isinstance(v, list) else v) for k, v in draws.items()},
                                      max_num_point=max_pts, full_scale=(128, 512))
    if max_pts < n:
        assert want["point_xyz"].shape[0] < n  # the crop engaged
    assert got["point_xyz"].shape == want["point_xyz"].shape
    assert np.abs(got["point_xyz"].cpu().numpy() - want["point_xyz"]).max() < 1e-6
    e_got, e_want = got["point_xyz_elastic"].cpu().numpy(), want["point_xyz_elastic"]
    assert np.abs(e_got - e_want).max() <= 2e-6
    assert np.array_equal(got["sem_labels"].cpu().numpy(), want["sem_labels"])
    assert np.array_equal(got["instance_ids"].cpu().numpy(), want["instance_ids"])
    assert int(got["num_instance"]) == int(want["num_instance"])
    assert np.array_equal(got["instance_num_point"].cpu().numpy(), want["instance_num_point"])
    assert np.array_equal(got["instance_semantic_cls"].cpu().numpy(), want["instance_semantic_cls"])
    fg = want["instance_ids"] >= 0
    assert np.abs(got["instance_center_xyz"].cpu().numpy()[fg] - want["instance_center_xyz"][fg]).max() < 1e-5
    # voxelisation: per-point voxel coordinates through the inverse map
    vx_g = got["voxel_xyz"].cpu().numpy()[got["voxel_point_map"].cpu().numpy()]
    vx_w = want["voxel_xyz"][want["voxel_point_map"]]
    differ = (vx_g != vx_w).any(1).mean()
    assert differ <= 5e-4, differ
    # the voxel set and the first-occurrence structure are self-consistent: every voxel's features are those of its
    # first point, every point maps to the voxel with its own quantised coordinate
    vmap = got["voxel_point_map"]
    q = torch.floor(got["point_xyz_elastic"] / 0.02).to(torch.int32)
    assert torch.equal(got["voxel_xyz"][vmap], q)
    first = torch.full((got["voxel_xyz"].size(0),), vmap.numel(), dtype=torch.int64, device="cuda").scatter_reduce_(
        0, vmap, torch.arange(vmap.numel(), device="cuda"), reduce="amin")
    assert torch.equal(first, torch.sort(first).values)  # first-occurrence order
    if differ == 0:  # identical quantisation -> identical voxel list, maps and features
        assert np.array_equal(got["voxel_xyz"].cpu().numpy(), want["voxel_xyz"])
        assert np.array_equal(got["voxel_point_map"].cpu().numpy(), want["voxel_point_map"])
        assert np.abs(got["voxel_features"].cpu().numpy() - want["voxel_features"]).max() < 1e-6


def test_elastic_kernels_match_scipy():
    """b2s_elastic_blur + b2s_elastic_apply against scipy.ndimage.convolve + RegularGridInterpolator directly."""
    from minsu3d_b200.harness import gpu_loader
    rng = np.random.default_rng(0)
    x = (rng.uniform(-150, 150, (50_000, 3))).astype(np.float32)
    np.random.seed(5)
    draws = []
    want = dataset_ref.elastic(x, 6.0, 40.0, draws)
    xg = torch.from_numpy(x).cuda().double()
    gpu_loader.elastic_gpu(xg, draws[0], 6.0, 40.0)
    assert np.abs(xg.cpu().numpy() - want).max() < 1e-9 * 200
