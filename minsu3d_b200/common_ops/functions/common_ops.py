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
