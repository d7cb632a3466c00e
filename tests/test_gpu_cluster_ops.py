"""GPU parity for the COMMON_OPS surface: ball query, BFS clustering, hierarchical aggregation,
segmented reductions, RoI pooling, IoU, mask labels.

Three-way: libb2s (CUDA, through the C ABI and the COMMON_OPS drop-in)  vs  the CPU oracle
(oracle/oracle.c)  vs  the reference's own compiled extension (oracle/_ref) when it is present.
Integer / index outputs must be bit-exact; the sequentially-defined fp32 sums are bit-exact too.
"""
import numpy as np
import pytest
import torch

import oracle
from helpers import canon_clusters, clustered_points

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _per_point_lists(idx, start_len):
    return [idx[s:s + l] for s, l in start_len]


# ---- C1 ball query ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,collapse,radius", [(0, False, 0.03), (1, False, 0.03), (3000, False, 0.03),
                                               (40_000, False, 0.03), (24_000, True, 0.03), (20_000, False, 0.04)])
def test_ballquery_matches_oracle(n, collapse, radius):
    from minsu3d_b200 import ops
    rng = np.random.default_rng(n + 7)
    if n >= 100:
        xyz, lab, bidx, offs = clustered_points(rng, n, collapse=collapse)
    else:
        xyz = rng.standard_normal((n, 3)).astype(np.float32)
        bidx = np.zeros(n, np.uint8)
        offs = np.array([0, n], np.int32)
    idx, sl = oracle.ballquery(xyz, bidx, offs, radius)
    g_idx, g_sl = ops.ballquery(_dev(xyz), _dev(bidx), _dev(offs), radius)
    assert np.array_equal(g_sl.cpu().numpy(), sl)
    assert np.array_equal(g_idx.cpu().numpy(), idx)
    if collapse:
        assert sl[:, 1].max() == 1000  # the cap (bfs_cluster.cu:38-43) is exercised


def test_ballquery_dense_lists_with_scattered_indices():
    """Long lists take the bitmap path only while their candidates span <= 65536 point indices; here two dense
    blobs have their members scattered over a 70k-point scene (merge fallback) and one is index-local (bitmap)."""
    from minsu3d_b200 import ops
    rng = np.random.default_rng(99)
    n = 70_000
    xyz = (rng.random((n, 3)) * 6.0).astype(np.float32)  # sparse background
    scattered = rng.choice(n, 1400, replace=False)
    xyz[scattered[:700]] = (np.array([1.0, 1.0, 1.0]) + rng.standard_normal((700, 3)) * 0.008).astype(np.float32)
    xyz[scattered[700:]] = (np.array([4.0, 2.0, 3.0]) + rng.standard_normal((700, 3)) * 0.02).astype(np.float32)
    xyz[50_000:51_200] = (np.array([2.0, 5.0, 1.0]) + rng.standard_normal((1200, 3)) * 0.006).astype(np.float32)
    bidx = np.zeros(n, np.uint8)
    offs = np.array([0, n], np.int32)
    idx, sl = oracle.ballquery(xyz, bidx, offs, 0.03)
    assert sl[:, 1].max() == 1000 and np.count_nonzero(sl[:, 1] >= 48) > 2000
    g_idx, g_sl = ops.ballquery(_dev(xyz), _dev(bidx), _dev(offs), 0.03)
    assert np.array_equal(g_sl.cpu().numpy(), sl)
    assert np.array_equal(g_idx.cpu().numpy(), idx)


def test_ballquery_vs_reference_kernel(ref_ops):
    """The reference's CUDA kernel on the same GPU: identical per-point neighbour lists."""
    from minsu3d_b200 import ops
    rng = np.random.default_rng(21)
    xyz, lab, bidx, offs = clustered_points(rng, 30_000)
    n = xyz.shape[0]
    d_xyz, d_b, d_o = _dev(xyz), _dev(bidx), _dev(offs)
    mean_active = 400
    r_idx = torch.zeros(n * mean_active, dtype=torch.int32, device="cuda")
    r_sl = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    n_active = ref_ops.ballquery_batch_p(d_xyz, d_b, d_o, r_idx, r_sl, n, mean_active, 0.03)
    assert n_active <= n * mean_active
    g_idx, g_sl = ops.ballquery(d_xyz, d_b, d_o, 0.03)
    assert g_idx.numel() == n_active
    ref_lists = _per_point_lists(r_idx.cpu().numpy(), r_sl.cpu().numpy())
    got_lists = _per_point_lists(g_idx.cpu().numpy(), g_sl.cpu().numpy())
    for a, b in zip(ref_lists, got_lists):
        assert np.array_equal(a, b)


def test_ballquery_dropin_retry_contract():
    """COMMON_OPS.ballquery_batch_p returns nActive > n*meanActive without writing past the buffer."""
    from minsu3d_b200 import COMMON_OPS
    rng = np.random.default_rng(2)
    xyz, lab, bidx, offs = clustered_points(rng, 6000, collapse=True)
    n = xyz.shape[0]
    idx = torch.zeros(n * 2, dtype=torch.int32, device="cuda")
    sl = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    n_active = COMMON_OPS.ballquery_batch_p(_dev(xyz), _dev(bidx), _dev(offs), idx, sl, n, 2, 0.03)
    assert n_active > n * 2 and int(idx.abs().sum()) == 0
    mean_active = n_active // n + 1
    idx = torch.zeros(n * mean_active, dtype=torch.int32, device="cuda")
    assert COMMON_OPS.ballquery_batch_p(_dev(xyz), _dev(bidx), _dev(offs), idx, sl, n, mean_active, 0.03) == n_active
    o_idx, o_sl = oracle.ballquery(xyz, bidx, offs, 0.03)
    assert np.array_equal(idx[:n_active].cpu().numpy(), o_idx)


# ---- C2 / C3 BFS clustering --------------------------------------------------------------------
def _graph(rng, n, collapse):
    xyz, lab, bidx, offs = clustered_points(rng, n, collapse=collapse)
    idx, sl = oracle.ballquery(xyz, bidx, offs, 0.03)
    return xyz, lab, bidx, offs, idx, sl


@pytest.mark.parametrize("n,collapse,thr", [(4000, False, 5), (30_000, False, 50), (24_000, True, 50)])
def test_pg_bfs_cluster_bit_exact(n, collapse, thr, request):
    from minsu3d_b200.common_ops.functions import pointgroup_ops
    rng = np.random.default_rng(n + 3)
    xyz, lab, bidx, offs, idx, sl = _graph(rng, n, collapse)
    lab = lab.copy()
    lab[rng.integers(0, lab.size, lab.size // 20)] = 7  # label noise splits components
    want_i, want_o = oracle.pg_bfs_cluster(lab, idx, sl, thr)
    got_i, got_o = pointgroup_ops.pg_bfs_cluster(_dev(lab), _dev(idx), _dev(sl), thr)
    assert np.array_equal(got_o.cpu().numpy(), want_o)
    assert np.array_equal(got_i.cpu().numpy(), want_i)  # incl. BFS visit order inside each cluster
    assert want_o.size > 1


def test_pg_bfs_cluster_vs_reference_cpu(ref_ops):
    from minsu3d_b200.common_ops.functions import pointgroup_ops
    rng = np.random.default_rng(77)
    xyz, lab, bidx, offs, idx, sl = _graph(rng, 20_000, False)
    ci = torch.empty(0, dtype=torch.int32)
    co = torch.empty(0, dtype=torch.int32)
    ref_ops.pg_bfs_cluster(torch.from_numpy(lab), torch.from_numpy(idx), torch.from_numpy(sl), ci, co, sl.shape[0], 20)
    got_i, got_o = pointgroup_ops.pg_bfs_cluster(_dev(lab), _dev(idx), _dev(sl), 20)
    assert torch.equal(got_o.cpu(), co) and torch.equal(got_i.cpu(), ci)
    # and the strict drop-in signature with CPU tensors (pointgroup.py:49-52)
    from minsu3d_b200 import COMMON_OPS
    di = torch.empty(0, dtype=torch.int32)
    do = torch.empty(0, dtype=torch.int32)
    COMMON_OPS.pg_bfs_cluster(torch.from_numpy(lab), torch.from_numpy(idx), torch.from_numpy(sl), di, do, sl.shape[0], 20)
    assert torch.equal(di, ci) and torch.equal(do, co)


@pytest.mark.parametrize("class_id,thr", [(3, 0.05), (0, 0.05)])
def test_sg_bfs_cluster_bit_exact(class_id, thr, ref_ops):
    from minsu3d_b200.common_ops.functions import softgroup_ops
    from minsu3d_b200.harness.scenes import POINT_NUM_AVG
    rng = np.random.default_rng(5 + class_id)
    xyz, lab, bidx, offs, idx, sl = _graph(rng, 30_000, class_id == 3)
    want_i, want_o = oracle.sg_bfs_cluster(POINT_NUM_AVG, idx, sl, thr, class_id)
    got_i, got_o = softgroup_ops.sg_bfs_cluster(POINT_NUM_AVG, _dev(idx), _dev(sl), thr, class_id)
    assert np.array_equal(got_o.cpu().numpy(), want_o) and np.array_equal(got_i.cpu().numpy(), want_i)
    ci = torch.empty(0, dtype=torch.int32)
    co = torch.empty(0, dtype=torch.int32)
    ref_ops.sg_bfs_cluster(torch.tensor(POINT_NUM_AVG, dtype=torch.float32), torch.from_numpy(idx), torch.from_numpy(sl),
                           ci, co, sl.shape[0], thr, class_id)
    assert torch.equal(got_o.cpu(), co) and torch.equal(got_i.cpu(), ci)


def test_cluster_empty_and_no_cluster():
    from minsu3d_b200.common_ops.functions import pointgroup_ops
    e_i, e_o = pointgroup_ops.pg_bfs_cluster(torch.zeros(0, dtype=torch.int16, device="cuda"),
                                             torch.zeros(0, dtype=torch.int32, device="cuda"),
                                             torch.zeros((0, 2), dtype=torch.int32, device="cuda"), 50)
    assert e_i.shape == (0, 2) and e_o.tolist() == [0]
    # isolated points: every component has size 1 < threshold
    n = 100
    sl = torch.stack((torch.arange(n), torch.ones(n, dtype=torch.long)), 1).int().cuda()
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    i, o = pointgroup_ops.pg_bfs_cluster(torch.zeros(n, dtype=torch.int16, device="cuda"), idx, sl, 2)
    assert i.shape == (0, 2) and o.tolist() == [0]


# ---- C4 hierarchical aggregation ---------------------------------------------------------------
@pytest.mark.parametrize("set_aggr", [False, True])
def test_hierarchical_aggregation(set_aggr, ref_ops):
    from minsu3d_b200.common_ops.functions import hais_ops
    from minsu3d_b200.harness.scenes import POINT_NUM_AVG, RADIUS_AVG
    rng = np.random.default_rng(31)
    xyz, lab, bidx, offs = clustered_points(rng, 36_000, n_obj=10, spread=0.05)
    # small satellites next to the objects become fragments
    sat = xyz[::40] + rng.normal(0, 0.02, xyz[::40].shape).astype(np.float32) + np.float32(0.12)
    xyz = np.concatenate((xyz, sat)).astype(np.float32)
    lab = np.concatenate((lab, lab[::40]))
    bidx = np.concatenate((bidx, bidx[::40]))
    order = np.argsort(bidx, kind="stable")
    xyz, lab, bidx = xyz[order], lab[order], bidx[order]
    offs = np.concatenate(([0], np.cumsum(np.bincount(bidx)))).astype(np.int32)
    idx, sl = oracle.ballquery(xyz, bidx, offs, 0.03)
    want_i, want_o = oracle.hierarchical_aggregation(lab, xyz, idx, sl, bidx, set_aggr, POINT_NUM_AVG, RADIUS_AVG)
    got_i, got_o = hais_ops.hierarchical_aggregation(_dev(lab), _dev(xyz), _dev(idx), _dev(sl), _dev(bidx), set_aggr,
                                                     POINT_NUM_AVG, RADIUS_AVG, -1)
    assert np.array_equal(got_o.cpu().numpy(), want_o)
    assert np.array_equal(got_i.cpu().numpy(), want_i)
    assert want_o.size > 2
    # the reference binary through its own Python wrapper semantics (hais_ops.py:8-73)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    outs = [torch.empty(0, dtype=torch.int32) for _ in range(2)] + [torch.empty(0)]
    fr_i, fr_o, fr_c = [torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int32), torch.empty(0)]
    k_i, k_o, k_c = [torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int32), torch.empty(0)]
    p_i, p_o, p_c = [torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int32), torch.empty(0)]
    pp_i, pp_o = torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.int32)
    ref_ops.hierarchical_aggregation(t(lab), t(xyz), t(bidx), t(idx), t(sl), fr_i, fr_o, fr_c, k_i, k_o, k_c, p_i, p_o, p_c,
                                     pp_i, pp_o, torch.tensor(POINT_NUM_AVG, dtype=torch.float32),
                                     torch.tensor(RADIUS_AVG, dtype=torch.float32), sl.shape[0], int(set_aggr), -1)
    if set_aggr:
        pp_i = pp_i[:pp_o[-1]]
        p_i, p_o = pp_i, pp_o
    ref_sets = canon_clusters(k_i.numpy(), k_o.numpy()) + canon_clusters(p_i.numpy(), p_o.numpy())
    got_sets = canon_clusters(got_i.cpu().numpy(), got_o.cpu().numpy())
    assert len(ref_sets) == len(got_sets)
    for a, b in zip(ref_sets, got_sets):  # absorbed-fragment order is racy in the reference: compare sets
        assert np.array_equal(a, b)
    if not set_aggr:
        assert np.array_equal(np.concatenate((k_i.numpy()[:, 1], p_i.numpy()[:, 1])), got_i.cpu().numpy()[:, 1])


# ---- S1-S3 segmented ops ---------------------------------------------------------------------
def _segments(rng, n_seg, max_len, with_empty=True):
    lens = rng.integers(1, max_len, n_seg)
    if with_empty and n_seg > 3:
        lens[2] = 0
    return np.concatenate(([0], np.cumsum(lens))).astype(np.int32)


@pytest.mark.parametrize("c", [3, 16, 32])
def test_segmented_reductions_bit_exact(c, ref_ops):
    from minsu3d_b200.common_ops.functions import common_ops, softgroup_ops
    rng = np.random.default_rng(c)
    offs = _segments(rng, 300, 4000, with_empty=False)
    x = rng.standard_normal((offs[-1], c)).astype(np.float32)
    dx, doffs = _dev(x), _dev(offs)
    for kind, fn in (("mean", common_ops.sec_mean), ("min", common_ops.sec_min), ("max", common_ops.sec_max)):
        got = fn(dx, doffs).cpu().numpy()
        assert np.array_equal(got, oracle.sec(kind, x, offs)), kind
        ref = torch.zeros_like(torch.from_numpy(got)).cuda()
        getattr(ref_ops, "sec_" + kind)(dx, doffs, ref, offs.size - 1, c)
        assert np.array_equal(got, ref.cpu().numpy()), kind + " vs reference kernel"
    got = softgroup_ops.global_avg_pool(dx, doffs).cpu().numpy()
    assert np.array_equal(got, oracle.gap_fp(x, offs))
    ref = torch.zeros_like(torch.from_numpy(got)).cuda()
    ref_ops.global_avg_pool_fp(dx, doffs, ref, offs.size - 1, c)
    assert np.array_equal(got, ref.cpu().numpy())


def test_roipool_forward_backward(ref_ops):
    from minsu3d_b200.common_ops.functions import common_ops, softgroup_ops
    rng = np.random.default_rng(8)
    c = 16
    offs = _segments(rng, 200, 3000, with_empty=False)
    x = np.round(rng.standard_normal((offs[-1], c)), 1).astype(np.float32)  # rounding creates ties: first max wins
    want, arg = oracle.roipool_fp(x, offs)
    xt = _dev(x).requires_grad_(True)
    y = common_ops.roipool(xt, _dev(offs))
    assert np.array_equal(y.detach().cpu().numpy(), want)
    g = rng.standard_normal(want.shape).astype(np.float32)
    y.backward(_dev(g))
    assert np.array_equal(xt.grad.cpu().numpy(), oracle.roipool_bp(x.shape[0], arg, g))
    r_out = torch.zeros(want.shape, device="cuda")
    r_arg = torch.zeros(want.shape, dtype=torch.int32, device="cuda")
    ref_ops.roipool_fp(_dev(x), _dev(offs), r_out, r_arg, offs.size - 1, c)
    assert np.array_equal(r_arg.cpu().numpy(), arg) and np.array_equal(r_out.cpu().numpy(), want)
    xt2 = _dev(x).requires_grad_(True)
    z = softgroup_ops.global_avg_pool(xt2, _dev(offs))
    z.backward(_dev(g))
    assert np.array_equal(xt2.grad.cpu().numpy(), oracle.gap_bp(x.shape[0], offs, g))


@pytest.mark.parametrize("n_seg,max_len", [(72, 24_000), (5, 60_000), (400, 5_000)])
def test_roipool_long_segments_multi_part_path(n_seg, max_len):
    """Proposals of ~10k points (the bench batch: 72 proposals over ~700k point entries) take the multi-CTA-per-segment
    path of b2s_roipool_fp_ws: values and FIRST arg-max (ties from rounding) bit-exact against the oracle, with an
    empty segment in the middle."""
    from minsu3d_b200.common_ops.functions import common_ops
    rng = np.random.default_rng(n_seg)
    offs = _segments(rng, n_seg, max_len, with_empty=True)
    x = np.round(rng.standard_normal((offs[-1], 16)), 1).astype(np.float32)
    want, arg = oracle.roipool_fp(x, offs)
    xt = _dev(x).requires_grad_(True)
    y = common_ops.roipool(xt, _dev(offs))
    keep = np.diff(offs) > 0  # the oracle / reference leave an empty segment's row at its initial value
    assert np.array_equal(y.detach().cpu().numpy()[keep], want[keep])
    g = rng.standard_normal(want.shape).astype(np.float32)
    g[~keep] = 0
    y.backward(_dev(g))
    assert np.array_equal(xt.grad.cpu().numpy(), oracle.roipool_bp(x.shape[0], arg, g))


def test_segmented_empty_segments():
    from minsu3d_b200.common_ops.functions import common_ops
    x = torch.arange(12, dtype=torch.float32, device="cuda").view(4, 3)
    offs = torch.tensor([0, 2, 2, 4], dtype=torch.int32, device="cuda")
    assert common_ops.sec_mean(x, offs)[1].tolist() == [0.0, 0.0, 0.0]
    assert torch.isinf(common_ops.sec_max(x, offs)[1]).all()
    assert common_ops.sec_mean(x, offs[:1]).shape == (0, 3)


# ---- I1 / I2 -----------------------------------------------------------------------------------
def _proposals(rng, n_points, n_prop, n_inst):
    inst = rng.integers(-1, n_inst, n_points).astype(np.int16)
    inst_num = np.bincount(inst[inst >= 0], minlength=n_inst).astype(np.int32)
    lens = rng.integers(50, 3000, n_prop)
    offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
    pidx = np.concatenate([np.sort(rng.choice(n_points, l, replace=False)) for l in lens]).astype(np.int32)
    # make proposals overlap instances strongly
    for p in range(n_prop):
        sel = np.nonzero(inst == (p % n_inst))[0]
        k = min(sel.size, lens[p] // 2)
        pidx[offs[p]:offs[p] + k] = sel[:k]
    cls = rng.integers(-1, 18, n_inst).astype(np.int16)
    return pidx, offs, inst, inst_num, cls


def test_iou_and_mask_label_bit_exact(ref_ops):
    from minsu3d_b200.common_ops.functions import common_ops
    rng = np.random.default_rng(4)
    pidx, offs, inst, inst_num, cls = _proposals(rng, 60_000, 150, 37)
    scores = rng.uniform(0, 1, pidx.size).astype(np.float32)
    d = [_dev(a) for a in (pidx, offs, inst, inst_num)]
    iou = common_ops.get_iou(*d)
    assert np.array_equal(iou.cpu().numpy(), oracle.get_iou(pidx, offs, inst, inst_num))
    assert torch.equal(iou, common_ops.get_mask_iou_on_cluster(*d))
    iou_p = common_ops.get_mask_iou_on_pred(*d, _dev(scores))
    assert np.array_equal(iou_p.cpu().numpy(), oracle.get_iou(pidx, offs, inst, inst_num, scores))
    r = torch.zeros_like(iou)
    ref_ops.get_iou(*d, r, inst_num.size, offs.size - 1)
    assert torch.equal(r, iou)
    r.zero_()
    ref_ops.get_mask_iou_on_pred(*d, r, inst_num.size, offs.size - 1, _dev(scores))
    assert torch.equal(r, iou_p)
    for thr in (0.5, 0.05):
        ml, mm = common_ops.get_mask_label(d[0], d[1], d[2], _dev(cls), d[3], iou, -1, thr)
        w_ml, w_mm = oracle.get_mask_label(pidx, offs, inst, cls, iou.cpu().numpy(), -1, thr)
        assert np.array_equal(ml.cpu().numpy(), w_ml) and np.array_equal(mm.cpu().numpy(), w_mm)
        r_ml = torch.zeros(pidx.size, dtype=torch.bool, device="cuda")
        r_mm = torch.zeros(pidx.size, dtype=torch.bool, device="cuda")
        ref_ops.get_mask_label(d[0], d[1], d[2], _dev(cls), iou, inst_num.size, offs.size - 1, -1, thr, r_ml, r_mm)
        assert torch.equal(r_ml, ml) and torch.equal(r_mm, mm)


# ---- clusters_voxelization ---------------------------------------------------
@pytest.mark.parametrize("n_cluster,scale,shape,idx_dtype", [(1, 50, 14, torch.int64), (37, 50, 14, torch.int64),
                                                             (200, 50, 14, torch.int32), (60, 5, 20, torch.int64)])
def test_clusters_voxelize_bit_exact(n_cluster, scale, shape, idx_dtype):
    """Fused kernel == the reference's torch expression sequence on the GPU == the C oracle, integer for integer."""
    from minsu3d_b200 import ops
    from minsu3d_b200.harness import models
    rng = np.random.default_rng(5 + n_cluster)
    n = 60_000
    coords = (rng.random((n, 3)) * 8.0).astype(np.float32)
    sizes = rng.integers(1, 3000, n_cluster)
    sizes[0] = 1  # a single-point cluster: extent 0 -> 1/0 = inf, clamped to `scale`
    offs = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    pts = np.concatenate([rng.integers(0, n, 1) + rng.integers(0, 400, s) % n for s in sizes]) % n  # local blobs
    cid = np.repeat(np.arange(n_cluster), sizes)
    idx = np.stack((cid, pts), 1).astype(np.int64)
    rand = rng.random((2, 3)).astype(np.float32)
    want = oracle.clusters_voxelize(idx, offs, coords, float(scale), shape, rand)
    d_idx = _dev(idx).to(idx_dtype)
    got = ops.clusters_voxelize(d_idx, _dev(offs), _dev(coords), scale, shape, _dev(rand))
    seq = models.clusters_voxel_coords_torch(_dev(idx), _dev(offs), _dev(coords), scale, shape, _dev(rand))
    assert np.array_equal(got.cpu().numpy(), seq.cpu().numpy())
    assert np.array_equal(got.cpu().numpy(), want)
    assert got[:, 1:].min() >= 0 and got[:, 1:].max() < shape + 1


def test_soft_grouping_all_classes_in_one_pass_matches_the_per_class_loop():
    """softgroup.py:43-86: the batched grouping (stacked classes, composite batch index, cluster_select mode 3) returns
    the proposals of the reference's per-class loop, point for point and in the same order."""
    from minsu3d_b200.harness import models
    cfg = models.Config.for_model("softgroup", proposal_source="gt_noise")
    rng = np.random.default_rng(31)
    xyz, lab, bidx, offs = clustered_points(rng, 40_000)
    n = xyz.shape[0]
    # soft scores: the blob label dominates, a second class often passes the 0.2 threshold as well
    scores = rng.random((n, cfg.classes)).astype(np.float32) * 0.15
    scores[np.arange(n), lab % cfg.classes] += 0.6
    second = (lab + 3) % cfg.classes
    scores[np.arange(n), second] += (rng.random(n) < 0.5) * 0.3
    scores /= scores.sum(1, keepdims=True)
    offsets = (rng.standard_normal((n, 3)) * 0.01).astype(np.float32)
    args = (cfg, _dev(scores), _dev(offsets), _dev(xyz), _dev(bidx))
    want = models.soft_grouping_loop(*args)
    got = models.soft_grouping(*args)
    assert want is not None and got is not None
    assert torch.equal(got[1], want[1])
    assert torch.equal(got[0], want[0])
    assert want[1].numel() - 1 >= 4  # several proposals, from more than one class


# ------------------------------------------------------------------------------------------
# instance post-processing on the device vs the oracle and the reference goldens
# ------------------------------------------------------------------------------------------
def _pp_check(got, want, conf_tol=1e-6):
    g = {k: v.cpu().numpy() for k, v in got.items()}
    assert np.array_equal(g["proposal"], np.asarray(want["proposal"])) if "proposal" in want else True
    assert np.array_equal(g["label_id"], np.asarray(want["label_id"]))
    assert np.allclose(g["conf"], np.asarray(want["conf"]), rtol=0, atol=conf_tol)
    assert np.array_equal(g["bbox"], np.asarray(want["bbox"]))
    assert np.array_equal(g["mask_offsets"], np.asarray(want["mask_offsets"]))
    assert np.array_equal(g["mask_points"], np.asarray(want["mask_points"]))


def _pp_inputs(c):
    import torch
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sem_scores = np.eye(20, dtype=np.float32)[np.asarray(c["semantic_labels"]).astype(np.int64)]
    return d(c["xyz"]), d(c["scores"]), d(c["proposals_idx"]), int(c["n_proposals"]), d(sem_scores), d(c["mask_scores"])


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_postprocess_matches_reference_goldens_and_oracle(ci):
    """PointGroup NMS path and HAIS filter path on the device = the reference's own methods (golden) = oracle."""
    import os
    from minsu3d_b200 import postprocess
    from oracle import postproc
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "postproc_ref.npz"))
    c = {k[len("c%d_in_" % ci):]: g[k] for k in g.files if k.startswith("c%d_in_" % ci)}
    xyz, scores, pidx, n_prop, sem_scores, mask_scores = _pp_inputs(c)
    args = (int(c["num_ignored"]), float(c["score_thr"]), int(c["npoint_thr"]))
    got = postprocess.pointgroup_pred_instances(xyz, scores, pidx, n_prop, sem_scores, *args, float(c["nms_thr"]))
    _pp_check(got, {k: g["c%d_pg_%s" % (ci, k)] for k in ("label_id", "conf", "bbox", "mask_offsets", "mask_points")})
    sem = np.asarray(c["semantic_labels"]).astype(np.int64)
    _pp_check(got, postproc.pointgroup_pred_instances(c["xyz"], c["scores"], c["proposals_idx"], n_prop, sem, *args,
                                                      float(c["nms_thr"])))
    got = postprocess.hais_pred_instances(xyz, scores, pidx, n_prop, mask_scores, sem_scores, int(c["num_ignored"]),
                                          float(c["mask_thr"]), float(c["score_thr"]), int(c["npoint_thr"]))
    _pp_check(got, {k: g["c%d_hais_%s" % (ci, k)] for k in ("label_id", "conf", "bbox", "mask_offsets", "mask_points")})
    from helpers import sg_mask_scores
    import torch
    dd = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sgm = sg_mask_scores(int(c["sg_seed"]), np.asarray(c["proposals_idx"]).shape[0])
    got_sg = postprocess.softgroup_pred_instances(xyz, pidx, xyz.size(0), dd(c["sg_cls_scores"]), dd(c["sg_iou_scores"]),
                                                  dd(sgm), int(c["instance_classes"]), float(c["mask_thr"]),
                                                  float(c["cls_thr"]), int(c["npoint_thr"]))
    _pp_check(got_sg, {k: g["c%d_sg_%s" % (ci, k)] for k in ("label_id", "conf", "bbox", "mask_offsets", "mask_points")})
    # the reference's list-of-dicts format incl. RLE strings round-trips to the same masks
    ref_fmt = postprocess.to_reference_format(got, "scene0000_00", int(np.asarray(c["xyz"]).shape[0]))
    assert len(ref_fmt) == got["label_id"].numel()
    if ref_fmt:
        runs = np.array(ref_fmt[0]["pred_mask"]["counts"].split(), np.int64)
        assert runs[1::2].sum() == int(got["mask_offsets"][1])


def test_postprocess_large_duplicates_and_empty():
    """200k points / 400 overlapping proposals with duplicated pairs (the dense mask is a set) vs the oracle;
    integer intermediates (npoint, intersections) bit-exact, IoU the same fp32 expression; nothing passes -> empty."""
    import torch
    from minsu3d_b200 import postprocess
    from oracle import postproc
    rng = np.random.default_rng(7)
    n, n_prop = 200_000, 400
    centers = rng.integers(0, n, n_prop)
    rows = []
    for p in range(n_prop):
        size = int(rng.integers(20, 3000))
        pts = (centers[p // 2 * 2] + rng.integers(-2000, 2000, size)) % n  # pairs of proposals share a neighbourhood
        rows.append(np.stack((np.full(size, p), pts), 1))  # rng.integers repeats points: duplicated pairs
    pidx = np.concatenate(rows).astype(np.int32)
    xyz = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    scores = rng.normal(0, 2, (n_prop, 1)).astype(np.float32)
    sem = rng.integers(0, 20, n)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    keys = postprocess._sorted_keys(d(pidx))
    mask = np.zeros((n_prop, n), bool)
    mask[pidx[:, 0], pidx[:, 1]] = True
    npoint = postprocess.proposal_npoint(keys, n_prop)
    assert np.array_equal(npoint.cpu().numpy(), mask.sum(1))
    keep = mask.sum(1) > 100
    ids, inter, iou = postprocess.proposal_cross_iou(keys, d(keep))
    mf = mask[keep].astype(np.float32)
    want_inter = mf @ mf.T
    assert np.array_equal(inter.cpu().numpy(), want_inter.astype(np.int32))
    npf = mf.sum(1)
    assert np.array_equal(iou.cpu().numpy(), want_inter / (npf[:, None] + npf[None, :] - want_inter))
    sem_scores = np.eye(20, dtype=np.float32)[sem]
    got = postprocess.pointgroup_pred_instances(d(xyz), d(scores), d(pidx), n_prop, d(sem_scores), 2, 0.09, 100, 0.3)
    want = postproc.pointgroup_pred_instances(xyz, scores, pidx, n_prop, sem, 2, 0.09, 100, 0.3)
    assert want["label_id"].size > 20
    _pp_check(got, want)
    got = postprocess.pointgroup_pred_instances(d(xyz), d(scores), d(pidx), n_prop, d(sem_scores), 2, 0.09, 10**6, 0.3)
    assert got["label_id"].numel() == 0 and got["bbox"].shape == (0, 6) and got["mask_offsets"].tolist() == [0]
