"""Instance post-processing on the device.

Replaces `PointGroup._get_pred_instances` + `_get_nms_instances` (minsu3d/model/pointgroup.py:197-265),
`HAIS._get_pred_instances` (minsu3d/model/hais.py:210-247) and `SoftGroup._get_pred_instances`
(minsu3d/model/softgroup.py:269-313): the reference moves everything to the CPU, builds dense
bool masks [nProposal, N], multiplies them (torch.mm) for the cross IoUs and runs a numpy NMS loop.  Here the
(proposal, point) pair list is sorted once on the GPU (csrc/postproc.cu) and distinct-point counts, the
intersection matrix, the IoUs and the NMS come out of it; labels and boxes are segmented reductions.

Both functions return the same dict of device tensors (order = the reference's instance order):
  proposal [n] int32, label_id [n] int64, conf [n] float32, bbox [n,6] float32,
  mask_points [sum] int32 (ascending inside an instance), mask_offsets [n+1] int32.
`to_reference_format` turns it into the reference's list of dicts (with its RLE strings) for the evaluation code.
"""
import numpy as np
import torch

from . import ops
from ._cabi import check, lib, ptr, require, require_cuda, stream, workspace

I32 = torch.int32


def _sorted_keys(proposals_idx, valid=None):
    require_cuda(proposals_idx)
    require(proposals_idx.dim() == 2 and proposals_idx.size(1) == 2, "proposals_idx must be [S,2] (proposal, point)")
    pidx = proposals_idx.to(I32).contiguous()
    S = pidx.size(0)
    dev = pidx.device
    keys = torch.empty(S, dtype=torch.int64, device=dev)
    v = None
    if valid is not None:
        v = valid.to(torch.uint8).contiguous()
        require(v.numel() == S, "valid must have one entry per pair")
    ws = workspace(lib().b2s_proposal_sort_ws_bytes(S), dev)
    check(lib().b2s_proposal_sort(ptr(pidx), ptr(v), S, ptr(keys), ptr(ws), ws.numel(), stream()), "proposal_sort")
    return keys


def proposal_npoint(keys, num_proposals):
    npoint = torch.empty(num_proposals, dtype=I32, device=keys.device)
    check(lib().b2s_proposal_npoint(ptr(keys), keys.numel(), int(num_proposals), ptr(npoint), stream()),
          "proposal_npoint")
    return npoint


def proposal_cross_iou(keys, keep):
    """keep: bool [nProposal].  Returns (ids int64 [n_kept], inter int32 [n_kept,n_kept], iou float32 [n_kept,n_kept])."""
    ids = torch.nonzero(keep).view(-1)
    n_kept = ids.numel()
    dev = keys.device
    remap = torch.full((keep.numel(),), -1, dtype=I32, device=dev)
    remap[ids] = torch.arange(n_kept, dtype=I32, device=dev)
    inter = torch.empty((n_kept, n_kept), dtype=I32, device=dev)
    iou = torch.empty((n_kept, n_kept), dtype=torch.float32, device=dev)
    check(lib().b2s_proposal_iou(ptr(keys), keys.numel(), ptr(remap), n_kept, ptr(inter), ptr(iou), stream()),
          "proposal_iou")
    return ids, inter, iou


def nms(cross_ious, scores, threshold):
    """pointgroup.py:197-218 on the device: indices of the picked proposals in pick order (int64)."""
    require_cuda(cross_ious, scores)
    n = scores.numel()
    dev = scores.device
    require(cross_ious.dtype == torch.float32 and cross_ious.is_contiguous() and tuple(cross_ious.shape) == (n, n),
            "cross_ious must be contiguous float32 [n,n]")
    # ties keep the lower index first (the reference's numpy argsort is unstable)
    order = torch.argsort(scores.view(-1), descending=True, stable=True).to(I32)
    pick = torch.empty(n, dtype=I32, device=dev)
    d_count = torch.empty(1, dtype=I32, device=dev)
    ws = workspace(max(n, 1), dev)
    check(lib().b2s_nms(ptr(cross_ious), ptr(order), n, float(threshold), ptr(pick), ptr(d_count), ptr(ws), ws.numel(),
                        stream()), "nms")
    return pick[:int(d_count.item())].long()


def _instances(keys, ids, conf, xyz, sem_labels, num_ignored, num_proposals, label=None):
    """Labels, boxes and point lists of the proposals `ids` (in that order) from the sorted distinct pairs.
    label: fixed label id for every instance (SoftGroup) instead of the semantic label of the first point."""
    dev = keys.device
    n = ids.numel()
    valid = keys != -1
    k = keys[valid]
    k = k[torch.cat((torch.ones(1, dtype=torch.bool, device=dev), k[1:] != k[:-1]))] if k.numel() else k  # distinct
    point = k >> 32
    prop = k & 0xFFFFFFFF
    rank = torch.full((num_proposals,), -1, dtype=torch.int64, device=dev)
    rank[ids] = torch.arange(n, device=dev)
    r = rank[prop]
    sel = r >= 0
    r, point = r[sel], point[sel]
    # keys are sorted by point: a stable sort by instance rank leaves the points ascending inside an instance
    order = torch.argsort(r, stable=True)
    r, point = r[order], point[order]
    counts = torch.bincount(r, minlength=n)
    offsets = torch.zeros(n + 1, dtype=I32, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    out = {"proposal": ids.to(I32), "conf": conf, "mask_points": point.to(I32), "mask_offsets": offsets}
    if n == 0:
        out["label_id"] = torch.zeros(0, dtype=torch.int64, device=dev)
        out["bbox"] = torch.zeros((0, 6), dtype=torch.float32, device=dev)
        return out
    if label is not None:
        out["label_id"] = (label.to(torch.int64) if torch.is_tensor(label)
                           else torch.full((n,), int(label), dtype=torch.int64, device=dev))
    else:
        first = point[offsets[:-1].long()]  # lowest point index of each instance (semantic_pred_labels[mask][0])
        out["label_id"] = sem_labels[first].long() - num_ignored + 1
    pts = xyz[point].contiguous()
    lo = torch.empty((n, 3), dtype=torch.float32, device=dev)
    hi = torch.empty((n, 3), dtype=torch.float32, device=dev)
    ops.sec_reduce("min", pts, offsets, lo)
    ops.sec_reduce("max", pts, offsets, hi)
    out["bbox"] = torch.cat((lo, hi), dim=1)
    return out


def pointgroup_pred_instances(xyz, proposals_scores, proposals_idx, num_proposals, semantic_scores, num_ignored,
                              score_thr, npoint_thr, nms_thr):
    """pointgroup.py:220-265; every argument a device tensor or a Python number."""
    sem_labels = semantic_scores.max(1)[1]
    score = torch.sigmoid(proposals_scores.view(-1))
    keys = _sorted_keys(proposals_idx)
    npoint = proposal_npoint(keys, num_proposals)
    keep = (score > score_thr) & (npoint > npoint_thr)
    ids, _, iou = proposal_cross_iou(keys, keep)
    kept_score = score[ids]
    pick = nms(iou, kept_score, nms_thr) if ids.numel() else ids
    return _instances(keys, ids[pick], kept_score[pick], xyz, sem_labels, num_ignored, num_proposals)


def hais_pred_instances(xyz, scores, proposals_idx, num_proposals, mask_scores, semantic_scores, num_ignored,
                        mask_thr, score_thr, npoint_thr):
    """hais.py:210-247."""
    sem_labels = semantic_scores.max(1)[1]
    score = torch.sigmoid(scores.view(-1))
    keys = _sorted_keys(proposals_idx, valid=mask_scores.view(-1) > mask_thr)
    npoint = proposal_npoint(keys, num_proposals)
    ids = torch.nonzero((score > score_thr) & (npoint >= npoint_thr)).view(-1)
    return _instances(keys, ids, score[ids], xyz, sem_labels, num_ignored, num_proposals)


def softgroup_pred_instances(xyz, proposals_idx, num_points, cls_scores, iou_scores, mask_scores, instance_classes,
                             mask_thr, cls_thr, min_npoint):
    """softgroup.py:269-313.  The reference filters once per instance class (a Python loop over 18 classes with dense
    [nProposal, N] masks); here every (class, proposal) candidate is one VIRTUAL proposal vp = class * nProposal +
    proposal, so one pair sort / distinct-point count / instance extraction serves all classes (three host reads per
    scene instead of ~110).  Output ordered by class, then proposal (= ascending vp); label_id = class + 1."""
    num_instances = cls_scores.size(0)
    dev = cls_scores.device
    cls = cls_scores.softmax(1)
    above = cls[:, :instance_classes] > cls_thr                                  # [P, C] candidates
    prop = proposals_idx[:, 0].long()
    # pair r is a point of candidate (class i, proposal prop[r]) iff the class-i mask keeps it and the candidate exists
    member = (mask_scores[:, :instance_classes] > mask_thr) & above[prop]        # [S, C]
    rows, klass = torch.nonzero(member, as_tuple=True)                           # host read 1: number of virtual pairs
    vpairs = torch.stack((klass * num_instances + prop[rows], proposals_idx[rows, 1].long()), dim=1)
    n_virtual = instance_classes * num_instances
    keys = _sorted_keys(vpairs) if vpairs.size(0) else torch.empty(0, dtype=torch.int64, device=dev)
    npoint = proposal_npoint(keys, n_virtual) if vpairs.size(0) else torch.zeros(n_virtual, dtype=I32, device=dev)
    ids = torch.nonzero(above.t().reshape(-1) & (npoint >= min_npoint)).view(-1)   # host read 2: instances (class-major)
    if ids.numel() == 0:
        return {"proposal": torch.zeros(0, dtype=I32, device=dev), "label_id": torch.zeros(0, dtype=torch.int64, device=dev),
                "conf": torch.zeros(0, dtype=torch.float32, device=dev),
                "bbox": torch.zeros((0, 6), dtype=torch.float32, device=dev),
                "mask_points": torch.zeros(0, dtype=I32, device=dev), "mask_offsets": torch.zeros(1, dtype=I32, device=dev)}
    klass_of, prop_of = ids // num_instances, ids % num_instances
    score = cls[prop_of, klass_of] * iou_scores[prop_of, klass_of].clamp(0, 1)
    out = _instances(keys, ids, score, xyz, None, 0, n_virtual, label=klass_of + 1)
    out["proposal"] = prop_of.to(I32)
    return out


def to_reference_format(result, scan_id, num_points):
    """The reference's list of dicts {scan_id, label_id, conf, pred_mask (RLE), pred_bbox} (pointgroup.py:255-264)."""
    pts = result["mask_points"].cpu().numpy()
    offs = result["mask_offsets"].cpu().numpy()
    label, conf, bbox = (result[k].cpu().numpy() for k in ("label_id", "conf", "bbox"))
    out = []
    for i in range(label.shape[0]):
        mask = np.zeros(num_points + 2, np.int8)
        mask[pts[offs[i]:offs[i + 1]] + 1] = 1
        runs = np.where(mask[1:] != mask[:-1])[0] + 1  # rle_encode, evaluation/instance_segmentation.py:10-23
        runs[1::2] -= runs[::2]
        out.append({"scan_id": scan_id, "label_id": int(label[i]), "conf": conf[i],
                    "pred_mask": {"length": num_points, "counts": " ".join(str(x) for x in runs)},
                    "pred_bbox": bbox[i]})
    return out
