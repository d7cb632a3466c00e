"""Host-side mirror of minsu3d/common_ops/functions/common_ops.py (same names, arguments, dtypes).

Differences that are deliberate and documented in DESIGN.md: the ball query sizes its output with
a count pass instead of the reference's allocate-n*meanActive-and-retry loop (common_ops.py:31-37),
so `meanActive` is accepted and ignored; outputs are allocated with torch.empty where every
element is written by the kernel.
"""
import torch
from torch.autograd import Function

from ... import ops


class BallQueryBatchP(Function):
    @staticmethod
    def forward(ctx, coords, batch_idxs, batch_offsets, radius, meanActive):
        idx, start_len = ops.ballquery(coords, batch_idxs, batch_offsets, radius)
        ctx.mark_non_differentiable(idx, start_len)
        return idx, start_len

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None


ballquery_batch_p = BallQueryBatchP.apply


def _sec(kind):
    class _Sec(Function):
        @staticmethod
        def forward(ctx, inp, offsets):
            out = torch.empty((offsets.size(0) - 1, inp.size(1)), dtype=torch.float32, device=inp.device)
            ops.sec_reduce(kind, inp, offsets, out)
            ctx.mark_non_differentiable(out)
            return out

        @staticmethod
        def backward(ctx, a=None):
            return None, None

    _Sec.__name__ = "Sec" + kind.capitalize()
    return _Sec


SecMean, SecMin, SecMax = _sec("mean"), _sec("min"), _sec("max")
sec_mean, sec_min, sec_max = SecMean.apply, SecMin.apply, SecMax.apply


class RoiPool(Function):
    @staticmethod
    def forward(ctx, feats, proposals_offset):
        n_prop = proposals_offset.size(0) - 1
        sum_npoint, c = feats.size()
        out = torch.empty((n_prop, c), dtype=torch.float32, device=feats.device)
        maxidx = torch.empty((n_prop, c), dtype=torch.int32, device=feats.device)
        ops.roipool_fp(feats, proposals_offset, out, maxidx)
        ctx.for_backwards = (maxidx, proposals_offset, sum_npoint)
        return out

    @staticmethod
    def backward(ctx, d_output_feats):
        maxidx, proposals_offset, sum_npoint = ctx.for_backwards
        c = d_output_feats.size(1)
        d_feats = torch.zeros((sum_npoint, c), dtype=torch.float32, device=d_output_feats.device)
        ops.roipool_bp(d_feats, proposals_offset, maxidx, d_output_feats.contiguous())
        return d_feats, None


roipool = RoiPool.apply


def _iou(proposals_idx, proposals_offset, instance_ids, instance_pointnum, mask_scores=None):
    n_inst = instance_pointnum.size(0)
    n_prop = proposals_offset.size(0) - 1
    iou = torch.empty((n_prop, n_inst), dtype=torch.float32, device=proposals_idx.device)
    assert proposals_idx.is_contiguous() and proposals_idx.is_cuda
    assert proposals_offset.is_contiguous() and proposals_offset.is_cuda
    assert instance_ids.is_contiguous() and instance_ids.is_cuda
    assert instance_pointnum.is_contiguous() and instance_pointnum.is_cuda
    ops.get_iou(proposals_idx, proposals_offset, instance_ids, instance_pointnum, iou, mask_scores)
    return iou


def get_iou(proposals_idx, proposals_offset, instance_ids, instance_pointnum):
    with torch.no_grad():
        return _iou(proposals_idx, proposals_offset, instance_ids, instance_pointnum)


get_mask_iou_on_cluster = get_iou


def get_mask_iou_on_pred(proposals_idx, proposals_offset, instance_labels, instance_pointnum, mask_scores_sigmoid):
    with torch.no_grad():
        assert mask_scores_sigmoid.is_contiguous() and mask_scores_sigmoid.is_cuda
        return _iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum,
                    mask_scores_sigmoid.detach().reshape(-1))


def get_mask_label(proposals_idx, proposals_offset, instance_ids, instance_cls, instance_pointnum, proposals_iou,
                   ignored_label, iou_thr):
    with torch.no_grad():
        n_inst = instance_pointnum.size(0)
        n_prop = proposals_offset.size(0) - 1
        mask_label = torch.zeros(proposals_idx.shape, dtype=torch.bool, device=proposals_idx.device)
        mask_label_mask = torch.zeros(proposals_idx.shape, dtype=torch.bool, device=proposals_idx.device)
        assert proposals_iou.is_contiguous() and proposals_iou.is_cuda
        ops.get_mask_label(proposals_idx, proposals_offset, instance_ids, instance_cls, proposals_iou, n_inst,
                           n_prop, ignored_label, iou_thr, mask_label, mask_label_mask)
        return mask_label, mask_label_mask


# ---------------------------------------------------------------------------------------------------------------
# voxelization_idx / voxelization: the names BASELINE.json's north_star uses for rows V1 / V2 of SURVEY.md 8(a).
# They do not exist in the reference (minsu3d replaced the upstream PointGroup ops by
# ME.utils.sparse_quantize(return_index, return_inverse), general_dataset.py:159-163 / general_model.py:187-189, and a
# plain gather features[v2p_map], backbone.py:40), so they are thin aliases with exactly those semantics.
# ---------------------------------------------------------------------------------------------------------------
def voxelization_idx(coords, batchsize=None, mode=4):
    """coords [N, 4] integer (batch, x, y, z) on the GPU -> (voxel_coords [M, 4] int32, p2v_map [N] int64,
    v2p_map [M] int64): unique rows in first-occurrence order (the libb2s coordinate hash, row V1).
    p2v_map[i] = voxel of point i (ME's inverse map); v2p_map[v] = first point of voxel v (ME's unique map).
    `batchsize` / `mode` are accepted for signature compatibility with the upstream op and not needed."""
    if not coords.is_cuda:
        raise ValueError("voxelization_idx needs CUDA coordinates (libb2s has no CPU path)")
    c = coords.to(torch.int32).contiguous()
    _, unique_idx, inverse, out_coords = ops.coord_unique(c, 1)
    return out_coords, inverse.long(), unique_idx.long()


class Voxelization(Function):
    """feats [N, C] -> voxel feats [M, C].  mode 4 = mean over the points of a voxel (upstream PointGroup default;
    scatter-add with vector atomics + count), any other mode = first point of the voxel (what ME's
    RANDOM_SUBSAMPLE quantisation, the reference's actual behaviour, keeps)."""

    @staticmethod
    def forward(ctx, feats, p2v_map, n_voxels, mode=4, v2p_map=None):
        feats = feats.contiguous()
        p2v_map = p2v_map.long().contiguous()
        ctx.mode = int(mode)
        if ctx.mode == 4:
            out = torch.zeros((n_voxels, feats.size(1)), dtype=torch.float32, device=feats.device)
            ops.check(ops.lib().b2s_scatter_add_rows(ops.ptr(feats), ops.ptr(p2v_map), feats.size(0), feats.size(1),
                                                     ops.ptr(out), ops.stream()), "scatter_add_rows")
            count = torch.bincount(p2v_map, minlength=n_voxels).clamp_(min=1).to(torch.float32)
            ctx.save_for_backward(p2v_map, count)
            return out / count[:, None]
        if v2p_map is None:
            raise ValueError("voxelization(mode != 4) needs v2p_map (first point of every voxel)")
        ctx.save_for_backward(v2p_map.long().contiguous())
        ctx.n = feats.size(0)
        return ops.devoxelize(feats, v2p_map.long().contiguous())

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        if ctx.mode == 4:
            p2v_map, count = ctx.saved_tensors
            return ops.devoxelize(grad / count[:, None], p2v_map), None, None, None, None
        (v2p_map,) = ctx.saved_tensors
        g = torch.zeros((ctx.n, grad.size(1)), dtype=torch.float32, device=grad.device)
        ops.check(ops.lib().b2s_scatter_add_rows(ops.ptr(grad), ops.ptr(v2p_map), grad.size(0), grad.size(1),
                                                 ops.ptr(g), ops.stream()), "scatter_add_rows")
        return g, None, None, None, None


voxelization = Voxelization.apply


def devoxelization(voxel_feats, p2v_map):
    """voxel feats [M, C] -> point feats [N, C] = voxel_feats[p2v_map] (backbone.py:40), scatter-add gradient."""
    return ops.devoxelize(voxel_feats.contiguous(), p2v_map.long().contiguous())
