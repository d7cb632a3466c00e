"""CPU suite (`-m "not gpu"`): the oracle against the reference's golden vectors and against independent
formulations, the C ABI surface, and the host-side logic.  No CUDA compute calls.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import oracle
from helpers import clustered_points, random_voxels, surface_voxels
from oracle import me_ref, me_unet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "common_ops_cpu.npz")
POINT_NUM_AVG = [-1, -1, 3917, 12056, 2303, 8331, 3948, 3166, 5629, 11719, 1003, 3317, 4912, 10221, 3889, 4136,
                 2120, 945, 3967, 2589]
RADIUS_AVG = [-1., -1., 0.7047687683952325, 1.1732690381942337, 0.39644035821116036, 1.011516629020215,
              0.7260155292902369, 0.8674973999335017, 0.8374931435447094, 1.0454153869133096, 0.32879464797430913,
              1.1954566226966346, 0.8628817944400078, 1.0416287916782507, 0.6602697958671507, 0.8541363897836871,
              0.38055290598206537, 0.3011878752684007, 0.7420871812436316, 0.4474268644407741]


# ------------------------------------------------------------------------------------------------
# oracle vs golden vectors produced by the reference's own compiled COMMON_OPS (tests/golden/make_golden.py)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_oracle_kat_from_reference_binary(golden):
    """6-point graph of SURVEY.md section 8(c), output computed by the reference binary."""
    ci, co = oracle.pg_bfs_cluster(golden["kat_lab"], golden["kat_idx"], golden["kat_sl"], 2)
    assert ci.tolist() == [[0, 0], [0, 1], [0, 2], [1, 3], [1, 4]] == golden["kat_ci"].tolist()
    assert co.tolist() == [0, 3, 5] == golden["kat_co"].tolist()


@pytest.mark.parametrize("tag", ["sparse", "dense"])
def test_oracle_bfs_and_ha_match_reference_golden(golden, tag):
    g = {k[len(tag) + 1:]: golden[k] for k in golden.files if k.startswith(tag + "_")}
    # the CSR input itself: oracle ball query == stored (brute force) lists
    idx, sl = oracle.ballquery(g["xyz"], g["bidx"], g["offs"], 0.03)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(sl, g["sl"])
    if tag == "dense":
        assert sl[:, 1].max() == 1000  # truncated (asymmetric) lists are covered
    ci, co = oracle.pg_bfs_cluster(g["lab"], idx, sl, 20)
    assert np.array_equal(ci, g["pg_ci"]) and np.array_equal(co, g["pg_co"])
    ci, co = oracle.sg_bfs_cluster(POINT_NUM_AVG, idx, sl, 0.05, 10)
    assert np.array_equal(ci, g["sg_ci"]) and np.array_equal(co, g["sg_co"])
    for grp, name in ((1, "kept"), (2, "prim")):
        ci, co, cc = oracle.bfs_cluster(g["lab"], idx, sl, 2, point_num_avg=POINT_NUM_AVG, group=grp,
                                        coords=g["xyz"], batch_idxs=g["bidx"])
        assert np.array_equal(ci, g["ha_%s_i" % name]) and np.array_equal(co, g["ha_%s_o" % name])
        assert np.array_equal(cc, g["ha_%s_c" % name])  # centres: sequential fp32 sums, bit-exact


def test_oracle_matches_reference_binary_live(ref_ops):
    """When oracle/_ref is present, re-run the reference's CPU entry points on fresh seeded inputs."""
    rng = np.random.default_rng(99)
    xyz, lab, bidx, offs = clustered_points(rng, 5000, n_obj=5)
    idx, sl = oracle.ballquery(xyz, bidx, offs, 0.03)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    ci, co = torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int32)
    ref_ops.pg_bfs_cluster(t(lab), t(idx), t(sl), ci, co, sl.shape[0], 10)
    o_ci, o_co = oracle.pg_bfs_cluster(lab, idx, sl, 10)
    assert np.array_equal(ci.numpy(), o_ci) and np.array_equal(co.numpy(), o_co)


def test_oracle_matches_reference_cuda_kernel_goldens():
    """tests/golden/common_ops_gpu.npz = outputs of the reference's own CUDA kernels on a B200 (make_golden.py --gpu)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "common_ops_gpu.npz"))
    idx, sl = oracle.ballquery(g["bq_xyz"], g["bq_bidx"], g["bq_offs"], 0.03)
    assert np.array_equal(sl[:, 1], g["bq_len"]) and np.array_equal(idx, g["bq_idx"])
    for kind in ("mean", "min", "max"):
        assert np.array_equal(oracle.sec(kind, g["seg_x"], g["seg_offs"]), g["sec_" + kind]), kind
    out, arg = oracle.roipool_fp(g["seg_x"], g["seg_offs"])
    assert np.array_equal(out, g["roipool_out"]) and np.array_equal(arg, g["roipool_arg"])
    assert np.array_equal(oracle.gap_fp(g["seg_x"], g["seg_offs"]), g["gap"])
    iou = oracle.get_iou(g["iou_pidx"], g["iou_poff"], g["iou_inst"], g["iou_inst_num"])
    assert np.array_equal(iou, g["iou"])
    iou_p = oracle.get_iou(g["iou_pidx"], g["iou_poff"], g["iou_inst"], g["iou_inst_num"], g["iou_scores"])
    assert np.array_equal(iou_p, g["iou_pred"])
    ml, mm = oracle.get_mask_label(g["iou_pidx"], g["iou_poff"], g["iou_inst"], g["iou_cls"], iou, -1, 0.25)
    assert np.array_equal(ml, g["mask_label"]) and np.array_equal(mm, g["mask_label_mask"])


def test_oracle_clusters_voxelize_matches_gpu_golden():
    """tests/golden/clusters_voxelize_gpu.npz = the reference's torch expression sequence (general_model.py:154-184)
    run on a B200 with the reference's own sec_* kernels (make_golden.py --gpu)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "clusters_voxelize_gpu.npz"))
    got = oracle.clusters_voxelize(g["idx"], g["offs"], g["coords"], float(g["scale"]), int(g["shape"]), g["rand"])
    assert np.array_equal(got, g["batched_xyz"])
    assert got[:, 1:].min() >= 0 and got[:, 1:].max() <= int(g["shape"])


# ------------------------------------------------------------------------------------------------
# MinkowskiEngine restatement: C oracle vs dictionary restatement vs dense torch conv3d
# ------------------------------------------------------------------------------------------------
def test_me_oracle_unique_and_stride_match_dictionary_restatement():
    rng = np.random.default_rng(1)
    c = random_voxels(rng, 3000, extent=24, dup=0.3)
    ui, inv, oc = oracle.coord_unique(c, 1)
    um, im = me_ref.sparse_quantize(c)
    assert np.array_equal(ui, um) and np.array_equal(inv, im) and np.array_equal(oc, c[um])
    for stride in (2, 4):
        _, inv2, oc2 = oracle.coord_unique(oc, stride)
        want_c, want_inv = me_ref.stride_coords(oc, stride)
        assert np.array_equal(oc2, want_c) and np.array_equal(inv2, want_inv)


@pytest.mark.parametrize("ksize,dil", [(3, 1), (3, 2), (2, 1), (1, 1)])
def test_me_oracle_kernel_map_matches_dictionary_restatement(ksize, dil):
    rng = np.random.default_rng(ksize + dil)
    c = oracle.coord_unique(random_voxels(rng, 800, extent=12), 1)[2]
    c[:, 1:] *= dil
    out_c = oracle.coord_unique(c, 2 * dil)[2] if ksize == 2 else c
    nbr = oracle.kernel_map(c, out_c, ksize, dil)
    maps = me_ref.kernel_map(c, out_c, ksize, dil)
    for k, pairs in enumerate(maps):
        got = [(int(nbr[o, k]), int(o)) for o in np.nonzero(nbr[:, k] >= 0)[0]]
        assert got == sorted(pairs, key=lambda p: p[1])
    pin, pout, koff = oracle.pairs_from_nbr(nbr)
    assert koff[-1] == sum(len(p) for p in maps) and np.all(np.diff(koff) >= 0)


def test_me_oracle_conv_equals_dense_conv3d_and_dictionary():
    """Independent of any memory of ME internals except the x-fastest offset convention (appendix A.4)."""
    rng = np.random.default_rng(7)
    c = random_voxels(rng, 1500, extent=10, batch=2)
    c[:, 1:] = np.abs(c[:, 1:]) % 8
    c = oracle.coord_unique(c, 1)[2]
    n, cin, cout = c.shape[0], 5, 7
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = rng.standard_normal((27, cin, cout)).astype(np.float32)
    nbr = oracle.kernel_map(c, c, 3, 1)
    got = oracle.conv_fwd(x, w, nbr, n)
    want_dict = me_ref.conv_forward(x, w, me_ref.kernel_map(c, c, 3, 1), n)
    assert np.abs(got - want_dict).max() < 1e-4
    dense = torch.zeros(2, cin, 8, 8, 8, dtype=torch.float64)
    ct = torch.from_numpy(c).long()
    dense[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]] = torch.from_numpy(x).double()
    wd = torch.from_numpy(w).double().view(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)
    out = torch.nn.functional.conv3d(dense, wd, padding=1)
    want = out[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]].numpy()
    assert np.abs(got - want).max() < 1e-4
    # backward by definition (finite structure): gin = sum_k gout[O_k] W_k^T, gW_k = in[I_k]^T gout[O_k]
    g = rng.standard_normal((n, cout)).astype(np.float32)
    gin, gw = oracle.conv_bwd(x, w, g, nbr)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    y = torch.zeros(n, cout, dtype=torch.float64)
    for k in range(27):
        o = np.nonzero(nbr[:, k] >= 0)[0]
        y = y.index_add(0, torch.from_numpy(o), xt[torch.from_numpy(nbr[o, k]).long()] @ wt[k])
    y.backward(torch.from_numpy(g).double())
    assert np.abs(gin - xt.grad.numpy()).max() < 1e-4 and np.abs(gw - wt.grad.numpy()).max() < 1e-3


def test_me_oracle_transposed_conv_is_adjoint_structure_of_strided_conv():
    rng = np.random.default_rng(8)
    c = surface_voxels(rng, 2000)
    _, _, oc = oracle.coord_unique(c, 2)
    nbr_down = oracle.kernel_map(c, oc, 2, 1)
    assert (nbr_down >= 0).sum() == c.shape[0]
    xc = rng.standard_normal((oc.shape[0], 6)).astype(np.float32)
    w = rng.standard_normal((8, 6, 4)).astype(np.float32)
    got = oracle.convT_fwd(xc, w, nbr_down, c.shape[0])
    want = me_ref.conv_transpose_forward(xc, w, me_ref.kernel_map(c, oc, 2, 1), c.shape[0])
    assert np.abs(got - want).max() < 1e-4


def test_me_oracle_strided_and_transposed_conv_equal_dense_torch():
    """2^3 / stride-2 convolution and its transposed form vs dense torch conv3d(stride=2) / conv_transpose3d(stride=2):
    pins the even-kernel offset enumeration (not centred, x fastest, appendix A.4-6) independently of me_ref."""
    rng = np.random.default_rng(12)
    c = random_voxels(rng, 1200, extent=10, batch=2)
    c[:, 1:] = np.abs(c[:, 1:]) % 8
    c = oracle.coord_unique(c, 1)[2]
    _, _, oc = oracle.coord_unique(c, 2)
    n, m, cin, cout = c.shape[0], oc.shape[0], 5, 6
    nbr = oracle.kernel_map(c, oc, 2, 1)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = rng.standard_normal((8, cin, cout)).astype(np.float32)
    ct, ot = torch.from_numpy(c).long(), torch.from_numpy(oc).long()
    dense = torch.zeros(2, cin, 8, 8, 8, dtype=torch.float64)
    dense[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]] = torch.from_numpy(x).double()
    wd = torch.from_numpy(w).double().view(2, 2, 2, cin, cout).permute(4, 3, 0, 1, 2)  # [cout, cin, iz, iy, ix]
    down = torch.nn.functional.conv3d(dense, wd, stride=2)
    want = down[ot[:, 0], :, ot[:, 3] // 2, ot[:, 2] // 2, ot[:, 1] // 2].numpy()
    got = oracle.conv_fwd(x, w, nbr, m)
    assert np.abs(got - want).max() < 1e-4
    # transposed: coarse [m, cout] -> fine [n, cin] with kernel (8, cout, cin); every fine voxel gets one term
    xc = rng.standard_normal((m, cout)).astype(np.float32)
    wt = rng.standard_normal((8, cout, cin)).astype(np.float32)
    dense_c = torch.zeros(2, cout, 4, 4, 4, dtype=torch.float64)
    dense_c[ot[:, 0], :, ot[:, 3] // 2, ot[:, 2] // 2, ot[:, 1] // 2] = torch.from_numpy(xc).double()
    wtd = torch.from_numpy(wt).double().view(2, 2, 2, cout, cin).permute(3, 4, 0, 1, 2)  # [in=cout, out=cin, iz, iy, ix]
    up = torch.nn.functional.conv_transpose3d(dense_c, wtd, stride=2)
    want_up = up[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]].numpy()
    got_up = oracle.convT_fwd(xc, wt, nbr, n)
    assert np.abs(got_up - want_up).max() < 1e-4


def test_oracle_backbone_forward_shapes_and_determinism():
    from minsu3d_b200.harness import models, scenes
    torch.manual_seed(123)
    model = models.build_model(models.Config.for_model("pointgroup"))
    n_param = sum(p.numel() for p in model.parameters())
    assert n_param == 7_715_320  # 7,532,128 + 999 + 182,176 + 17
    sd = me_unet.numpy_state_dict(model)
    data = scenes.collate([scenes.make_scene(3, n_points=4000)], "cpu")
    args = (sd, data["voxel_features"].numpy(), data["voxel_xyz"].numpy(), data["voxel_point_map"].numpy())
    a = me_unet.backbone_forward(*args)
    b = me_unet.backbone_forward(*args)
    assert a["semantic_scores"].shape == (4000, 20) and a["point_offsets"].shape == (4000, 3)
    assert np.array_equal(a["point_features"], b["point_features"])


# ------------------------------------------------------------------------------------------------
# C ABI: the library loads and exports every symbol include/b2s.h declares
# ------------------------------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol():
    from minsu3d_b200 import _cabi
    from minsu3d_b200.csrc import build as b2s_build
    lib_path = b2s_build.build()
    with open(os.path.join(ROOT, "include", "b2s.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    handle = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(handle, name), "libb2s.so does not export %s" % name
    assert set(_cabi.EXPORTED_SYMBOLS) == declared, declared ^ set(_cabi.EXPORTED_SYMBOLS)
    assert handle.b2s_version() == 100
    nm = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b2s_[a-z0-9_]+)", nm))
    assert exported == declared, exported ^ declared


def test_cabi_size_queries_and_sass_is_blackwell_native():
    from minsu3d_b200 import _cabi
    lib = _cabi.lib()
    assert lib.b2s_hash_capacity(1000) == 2048 and lib.b2s_hash_capacity(0) == 1024
    for fn, args in (("b2s_coord_unique_ws_bytes", (100_000,)), ("b2s_pairs_ws_bytes", (100_000, 27)),
                     ("b2s_ballquery_ws_bytes", (100_000,)), ("b2s_cluster_ws_bytes", (100_000,)),
                     ("b2s_bn_ws_bytes", (100_000, 16)), ("b2s_conv_ws_bytes", (27, 64, 64))):
        assert getattr(lib, fn)(*args) > 0
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    res = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True)
    assert res.returncode == 0 and len(res.stdout) > 100_000, "cuobjdump produced no SASS: %s" % res.stderr[:200]
    # tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, cp.async -> LDGSTS (B200_PROFILING.md)
    # (+ HMMA: the mma.sync TF32 weight gradient / warp-stream convolution)
    counts = {m: res.stdout.count(m) for m in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "LDGSTS", "UTCBAR", "HMMA")}
    for mnemonic, n in counts.items():
        assert n > 0, "%s missing from the SASS of libb2s.so (%s)" % (mnemonic, counts)


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: CUDA-only entry points refuse CPU tensors."""
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((ValueError, RuntimeError)):
        ME.SparseTensor(features=torch.zeros(4, 3), coordinates=torch.zeros(4, 4, dtype=torch.int32))
    with pytest.raises((ValueError, RuntimeError)):
        ops.ballquery(torch.zeros(4, 3), torch.zeros(4, dtype=torch.uint8), torch.tensor([0, 4], dtype=torch.int32), 0.03)
    with pytest.raises(RuntimeError):
        import bench
        bench.run_train(type("A", (), {"steps": 1, "warmup": 1, "no_cpu_baseline": True, "gpus": 1, "overlap": False,
                                       "size_hints": False})(), "pointgroup")


# ------------------------------------------------------------------------------------------------
# host logic: quantize (DataLoader-worker API), collate, module surface, state-dict names
# ------------------------------------------------------------------------------------------------
def test_sparse_quantize_cpu_api_first_occurrence():
    from minsu3d_b200 import MinkowskiEngine as ME
    pts = np.array([[0.011, 0.0, 0.0], [0.05, 0.0, 0.0], [0.012, 0.013, 0.019], [0.059, 0.001, 0.0], [-0.01, 0, 0]])
    feats = np.arange(10, dtype=np.float32).reshape(5, 2)
    vx, vf, um, inv = ME.utils.sparse_quantize(pts, feats, return_index=True, return_inverse=True, quantization_size=0.02)
    assert vx.tolist() == [[0, 0, 0], [2, 0, 0], [-1, 0, 0]] and vx.dtype == np.int32
    assert um.tolist() == [0, 1, 4] and inv.tolist() == [0, 1, 0, 1, 2] and um.dtype == torch.int64
    assert np.array_equal(vf, feats[[0, 1, 4]])
    um2, inv2 = me_ref.sparse_quantize(pts, 0.02)
    assert np.array_equal(um.numpy(), um2) and np.array_equal(inv.numpy(), inv2)
    only = ME.utils.sparse_quantize(torch.tensor([[1, 2, 3], [1, 2, 3]]), return_maps_only=True)
    assert only.tolist() == [0]


def test_sparse_collate_and_scene_contract():
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200.harness import scenes
    c, f = ME.utils.sparse_collate([np.zeros((2, 3), np.int32), np.ones((3, 3), np.int32)],
                                   [np.zeros((2, 6), np.float32), np.ones((3, 6), np.float32)])
    assert c.dtype == torch.int32 and c[:, 0].tolist() == [0, 0, 1, 1, 1] and f.shape == (5, 6)
    a, b = scenes.make_scene(5, n_points=3000), scenes.make_scene(5, n_points=3000)
    assert np.array_equal(a["xyz"], b["xyz"])  # seeded
    d = scenes.collate([a, scenes.make_scene(6, n_points=3000)], "cpu")
    dtypes = {"point_xyz": torch.float32, "vert_batch_ids": torch.uint8, "sem_labels": torch.int16,
              "instance_ids": torch.int16, "instance_center_xyz": torch.float32, "instance_num_point": torch.int32,
              "instance_offsets": torch.int32, "instance_semantic_cls": torch.int16, "voxel_xyz": torch.int32,
              "voxel_features": torch.float32, "voxel_point_map": torch.int64}
    for k, dt in dtypes.items():
        assert d[k].dtype == dt, k
    assert d["voxel_xyz"].shape[1] == 4 and d["voxel_features"].shape[1] == 6
    assert int(d["voxel_point_map"].max()) == d["voxel_xyz"].shape[0] - 1
    assert int(d["instance_ids"].max()) + 1 == d["instance_num_point"].numel() == int(d["instance_offsets"][-1])


def test_module_surface_and_state_dict_names_match_reference():
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200.harness import models
    conv = ME.MinkowskiConvolution(6, 16, kernel_size=3, dimension=3)
    assert tuple(conv.kernel.shape) == (27, 6, 16) and conv.bias is None
    assert float(conv.kernel.abs().max()) <= 1 / np.sqrt(6 * 27) + 1e-6
    assert tuple(ME.MinkowskiConvolution(32, 16, kernel_size=1, dimension=3).kernel.shape) == (32, 16)
    up = ME.MinkowskiConvolutionTranspose(32, 16, kernel_size=2, stride=2, dimension=3)
    assert tuple(up.kernel.shape) == (8, 32, 16) and float(up.kernel.abs().max()) <= 1 / np.sqrt(16 * 8) + 1e-6
    assert list(ME.MinkowskiBatchNorm(16).state_dict()) == ["bn.weight", "bn.bias", "bn.running_mean", "bn.running_var",
                                                           "bn.num_batches_tracked"]
    with pytest.raises(NotImplementedError):
        ME.MinkowskiConvolution(4, 4, kernel_size=3, dimension=2)
    keys = set(models.build_model(models.Config.for_model("pointgroup")).state_dict())
    for k in ("backbone.unet.0.kernel", "backbone.unet.1.blocks.block0.conv_branch.0.bn.weight",
              "backbone.unet.1.blocks.block1.conv_branch.5.kernel", "backbone.unet.1.conv.2.kernel",
              "backbone.unet.1.u.u.u.u.u.u.blocks.block0.conv_branch.2.kernel", "backbone.unet.1.deconv.2.kernel",
              "backbone.unet.1.blocks_tail.block0.downsample.0.kernel", "backbone.unet.2.bn.running_mean",
              "backbone.semantic_branch.3.weight", "backbone.offset_branch.0.bias", "score_net.unet.0.conv.2.kernel",
              "score_branch.weight"):
        assert k in keys, k
    n32 = sum(p.numel() for p in models.build_model(models.Config.for_model("hais")).backbone.unet.parameters())
    assert n32 == 30_106_304  # SURVEY.md section 8(e), m = 32


def test_install_as_reference_modules():
    import minsu3d_b200
    me, co = minsu3d_b200.install_as_reference_modules()
    import COMMON_OPS
    import MinkowskiEngine
    assert MinkowskiEngine is me and COMMON_OPS is co
    reference_names = ["sg_bfs_cluster", "global_avg_pool_fp", "global_avg_pool_bp", "ballquery_batch_p", "sec_mean",
                       "sec_min", "sec_max", "roipool_fp", "roipool_bp", "get_iou", "get_mask_iou_on_cluster",
                       "get_mask_iou_on_pred", "get_mask_label", "pg_bfs_cluster", "hierarchical_aggregation"]
    for n in reference_names:  # minsu3d/common_ops/src/common_ops_api.cpp:6-29
        assert callable(getattr(COMMON_OPS, n)), n
    for n in ("SparseTensor", "MinkowskiConvolution", "MinkowskiConvolutionTranspose", "MinkowskiBatchNorm",
              "MinkowskiReLU", "cat"):
        assert hasattr(MinkowskiEngine, n), n
    assert hasattr(MinkowskiEngine.utils, "sparse_quantize") and hasattr(MinkowskiEngine.utils, "sparse_collate")


def test_segmented_scores_and_offset_loss_helpers():
    from minsu3d_b200.harness import models
    s = models.get_segmented_scores(torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9]), 0.75, 0.25)
    assert torch.allclose(s, torch.tensor([0.0, 0.0, 0.5, 1.0, 1.0]))
    n, d = models.pt_offset_loss(torch.ones(4, 3), torch.ones(4, 3), torch.tensor([True, True, False, True]))
    assert float(n) == 0.0 and abs(float(d) + 1.0) < 1e-6
    assert models.pt_offset_loss(torch.ones(2, 3), torch.ones(2, 3), torch.tensor([False, False])) == (0, 0)


def test_offset_loss_masked_form_equals_reference_indexing_form():
    """harness.models.pt_offset_loss is written as masked means (no data-dependent shapes on the device); the
    reference (minsu3d/loss/pt_offset_loss.py:12-38) indexes with the boolean mask.  Same values and gradients."""
    import torch.nn.functional as F
    from minsu3d_b200.harness import models
    g = torch.Generator().manual_seed(0)
    pred = torch.randn(5000, 3, generator=g, dtype=torch.float64, requires_grad=True)
    gt = torch.randn(5000, 3, generator=g, dtype=torch.float64)
    mask = torch.rand(5000, generator=g) > 0.4
    n_a, d_a = models.pt_offset_loss(pred, gt, mask)
    (n_a + d_a).backward()
    grad_a = pred.grad.clone()
    pred.grad = None
    p, q = pred[mask], gt[mask]
    n_b = torch.sum(torch.abs(p - q), dim=-1).mean()
    eps = torch.finfo(q.dtype).eps
    d_b = -(F.normalize(q, p=2, dim=1, eps=eps) * F.normalize(p, p=2, dim=1, eps=eps)).sum(-1).mean()
    (n_b + d_b).backward()
    assert abs(float(n_a) - float(n_b)) < 1e-12 and abs(float(d_a) - float(d_b)) < 1e-12
    assert torch.allclose(grad_a, pred.grad, rtol=0, atol=1e-14)


# ------------------------------------------------------------------------------------------
# post-processing restatement pinned by the reference's own methods
# ------------------------------------------------------------------------------------------
def _postproc_case(g, ci):
    return {k[len("c%d_in_" % ci):]: g[k] for k in g.files if k.startswith("c%d_in_" % ci)}


def _check_instances(got, g, prefix):
    assert np.array_equal(np.asarray(got["label_id"]), g[prefix + "label_id"])
    assert np.allclose(np.asarray(got["conf"]), g[prefix + "conf"], rtol=0, atol=1e-6)
    assert np.array_equal(np.asarray(got["bbox"]), g[prefix + "bbox"])
    assert np.array_equal(np.asarray(got["mask_offsets"]), g[prefix + "mask_offsets"])
    assert np.array_equal(np.asarray(got["mask_points"]), g[prefix + "mask_points"])


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_postproc_oracle_matches_reference_methods(ci):
    """oracle.postproc vs outputs of PointGroup._get_pred_instances/_get_nms_instances and
    HAIS._get_pred_instances run from /root/reference (tests/golden/make_postproc_golden.py)."""
    from oracle import postproc
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "postproc_ref.npz"))
    c = _postproc_case(g, ci)
    sem = c["semantic_labels"].astype(np.int64)
    pg = postproc.pointgroup_pred_instances(c["xyz"], c["scores"], c["proposals_idx"], int(c["n_proposals"]), sem,
                                            int(c["num_ignored"]), float(c["score_thr"]), int(c["npoint_thr"]),
                                            float(c["nms_thr"]))
    _check_instances(pg, g, "c%d_pg_" % ci)
    assert pg["label_id"].size > 0
    hs = postproc.hais_pred_instances(c["xyz"], c["scores"], c["proposals_idx"], int(c["n_proposals"]),
                                      c["mask_scores"], sem, int(c["num_ignored"]), float(c["mask_thr"]),
                                      float(c["score_thr"]), int(c["npoint_thr"]))
    _check_instances(hs, g, "c%d_hais_" % ci)
    from helpers import sg_mask_scores
    sg = postproc.softgroup_pred_instances(c["xyz"], c["proposals_idx"], c["xyz"].shape[0], c["sg_cls_scores"],
                                           c["sg_iou_scores"], sg_mask_scores(int(c["sg_seed"]), c["proposals_idx"].shape[0]),
                                           int(c["instance_classes"]), float(c["mask_thr"]), float(c["cls_thr"]),
                                           int(c["npoint_thr"]))
    _check_instances(sg, g, "c%d_sg_" % ci)
    assert sg["label_id"].size > 0


# ------------------------------------------------------------------------------------------
# checkpoint compatibility pinned by the reference's own module classes
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["pointgroup", "hais", "softgroup"])
def test_state_dict_names_and_shapes_equal_reference_classes(name):
    """tests/golden/reference_state_dicts.json = state_dict() of the reference's PointGroup / HAIS / SoftGroup
    (instantiated from /root/reference by tests/golden/make_state_dict_golden.py)."""
    import json
    from minsu3d_b200.harness import models
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_state_dicts.json")) as f:
        want = json.load(f)[name]
    model = models.build_model(models.Config.for_model(name))
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert set(got) == set(want)
    assert got == want


def test_load_reference_checkpoint_roundtrip_and_errors(tmp_path):
    from minsu3d_b200.harness import checkpoint, models
    torch.manual_seed(0)
    src = models.build_model(models.Config.for_model("pointgroup"))
    path = str(tmp_path / "PointGroup_best.ckpt")
    checkpoint.save_reference_checkpoint(src, path, epoch=3)
    assert set(torch.load(path, weights_only=False)) >= {"state_dict", "epoch"}
    torch.manual_seed(1)
    dst = models.build_model(models.Config.for_model("pointgroup"))
    assert checkpoint.load_reference_checkpoint(dst, path) == ([], [])
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(ValueError, match="shape mismatches"):  # a HAIS (m=32) checkpoint does not fit PointGroup (m=16)
        checkpoint.load_reference_checkpoint(dst, models.build_model(models.Config.for_model("hais")).state_dict())
    sd = src.state_dict()
    sd.pop("score_branch.weight")
    with pytest.raises(ValueError, match="1 missing"):
        checkpoint.load_reference_checkpoint(dst, {"state_dict": sd})
    assert checkpoint.load_reference_checkpoint(dst, {"state_dict": sd}, strict=False) == (["score_branch.weight"], [])
    with pytest.raises(ValueError, match="not a checkpoint"):
        checkpoint.load_reference_checkpoint(dst, {"state_dict": 3})


def test_deferred_count_checks_host_logic():
    """ops.defer_count_check / run_deferred_checks: claims are validated at the next host read; a wrong claim or a
    range error raises and clears the list (host logic only: plain CPU tensors stand in for the device counts)."""
    from minsu3d_b200 import ops
    ops.run_deferred_checks()  # empty list: no-op
    ops.defer_count_check(torch.tensor([10, 0], dtype=torch.int32), 10, "a")
    ops.defer_count_check(torch.tensor([7, 0], dtype=torch.int32), 7, "b")
    ops.run_deferred_checks()
    ops.defer_count_check(torch.tensor([10, 0], dtype=torch.int32), 10, "ok")
    ops.defer_count_check(torch.tensor([9, 0], dtype=torch.int32), 10, "level 4")
    with pytest.raises(ValueError, match="level 4: claimed 10 rows but the device counted 9"):
        ops.run_deferred_checks()
    ops.run_deferred_checks()  # cleared by the failure
    ops.defer_count_check(torch.tensor([5, 1], dtype=torch.int32), 5, "range")
    with pytest.raises(ValueError, match="packable range"):
        ops.run_deferred_checks()


def test_tile_order_restatement_properties():
    """oracle.tile_order (what tests/test_gpu_maps_conv.py checks the CUDA schedule against): a stable permutation
    grouped by (stencil faces, neighbour mask); tile masks are the OR of their rows; far fewer active offsets."""
    rng = np.random.default_rng(4)
    c = surface_voxels(rng, 20_000, batch=2)
    nbr = oracle.kernel_map(c, c, 3, 1)
    perm, nbr_sorted, tile_mask = oracle.tile_order(nbr)
    n = nbr.shape[0]
    assert np.array_equal(np.sort(perm), np.arange(n)) and np.array_equal(nbr_sorted, nbr[perm])
    masks = ((nbr_sorted >= 0).astype(np.int64) << np.arange(27)).sum(1)
    change = np.nonzero(masks[1:] != masks[:-1])[0] + 1
    for a, b in zip(np.concatenate(([0], change)), np.concatenate((change, [n]))):
        assert np.all(np.diff(perm[a:b]) > 0)  # stable: equal keys keep first-occurrence order
    assert len(np.unique(masks)) == len(change) + 1  # every mask forms one contiguous run
    for t in (0, len(tile_mask) // 2, len(tile_mask) - 1):
        want = np.bitwise_or.reduce(masks[t * 128:(t + 1) * 128])
        assert int(tile_mask[t]) == int(want)
    popc = lambda ms: np.mean([bin(int(m)).count("1") for m in ms])
    shuffled = ((nbr >= 0).reshape(-1, 27)[rng.permutation(n)][:n // 128 * 128].reshape(-1, 128, 27).any(1)).sum(1).mean()
    assert popc(tile_mask) < 0.6 * shuffled


def test_postproc_restatement_edge_cases():
    from oracle import postproc
    n = 50
    xyz = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    sem = np.full(n, 5)
    # two identical proposals (IoU 1) with tied scores, a disjoint one, duplicated pairs inside proposal 0
    p0 = np.arange(0, 20)
    pidx = np.concatenate([np.stack((np.full(20, 0), p0), 1), np.stack((np.full(5, 0), p0[:5]), 1),
                           np.stack((np.full(20, 1), p0), 1), np.stack((np.full(15, 2), np.arange(30, 45)), 1)]).astype(np.int32)
    scores = np.array([[1.0], [1.0], [0.5]], np.float32)
    out = postproc.pointgroup_pred_instances(xyz, scores, pidx, 3, sem, 2, 0.09, 10, 0.3)
    assert out["proposal"].tolist() == [0, 2]  # tie: the lower proposal wins, its twin is suppressed
    assert out["mask_offsets"].tolist() == [0, 20, 35]  # duplicated pairs count once (the dense mask is a set)
    assert out["label_id"].tolist() == [4, 4] and np.array_equal(out["bbox"][1], np.concatenate((xyz[30], xyz[44])))
    assert np.array_equal(postproc.nms(np.eye(3, dtype=np.float32), np.array([0.1, 0.9, 0.5], np.float32), 0.3), [1, 2, 0])
    empty = postproc.pointgroup_pred_instances(xyz, scores, pidx, 3, sem, 2, 0.09, 100, 0.3)
    assert empty["label_id"].size == 0 and empty["mask_offsets"].tolist() == [0] and empty["bbox"].shape == (0, 6)
    hs = postproc.hais_pred_instances(xyz, scores, pidx, 3, np.ones((pidx.shape[0], 1), np.float32), sem, 2, -0.5, 0.09, 15)
    assert hs["proposal"].tolist() == [0, 1, 2]  # no NMS in HAIS, >= on the point count


# ------------------------------------------------------------------------------------------------
# the numpy restatement of the loader's sample pipeline pinned by the reference's own transform code
# ------------------------------------------------------------------------------------------------
def test_dataset_restatement_matches_reference_transform_functions():
    """oracle/dataset_ref.py calls np.random in the reference's order: seeded equally, the reference's own jitter /
    flip / rotz / elastic (minsu3d/util/transform.py, imported from /root/reference or the staged copy) give
    bit-identical results."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ref_loader import load_reference_models
    if load_reference_models() is None:
        pytest.skip("reference python not available")
    from minsu3d.util import transform as T
    from minsu3d_b200.harness import scenes
    from oracle import dataset_ref
    sc = scenes.make_scene(3, 20000)
    np.random.seed(11)
    data, draws = dataset_ref.train_sample(sc)
    np.random.seed(11)
    m = np.eye(3)
    m = np.matmul(m, T.jitter())
    m *= T.flip(0, random=True)
    m = np.matmul(m, T.rotz(np.random.rand() * 2 * np.pi)).astype(np.float32)
    xyz = np.matmul(sc["xyz"].astype(np.float32), m)
    rgb_jit = np.random.randn(3) * 0.1
    scale = 1 / 0.02
    e = T.elastic(xyz * scale, 6 * scale // 50, 40 * scale / 50)
    e = T.elastic(e, 20 * scale // 50, 160 * scale / 50)
    e -= e.min(axis=0)
    e /= scale
    assert np.array_equal(m, draws["aug_matrix"]) and np.array_equal(rgb_jit, draws["rgb_jitter"])
    assert np.array_equal(xyz, data["point_xyz"])
    assert np.array_equal(e, data["point_xyz_elastic"])
    # crop with a small budget: same valid set as the reference's crop on the same draws
    np.random.seed(3)
    pc = e * scale
    a, va = dataset_ref.crop(pc, 8000, 512, [])
    np.random.seed(3)
    b, vb = T.crop(pc, 8000, 512)
    assert np.array_equal(va, vb) and np.array_equal(a, b) and va.sum() <= 8000
