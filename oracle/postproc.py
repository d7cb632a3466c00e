"""oracle.postproc -- numpy restatement of the reference's instance post-processing.  TEST INFRASTRUCTURE ONLY.

Follows minsu3d/model/pointgroup.py:197-265 (`_get_nms_instances`, `_get_pred_instances`) and
minsu3d/model/hais.py:210-247 (`_get_pred_instances`) step by step, dense [nProposal, N] masks included.
PINNED: tests/golden/postproc_ref.npz holds outputs of the reference's own methods (executed from
/root/reference by tests/golden/make_postproc_golden.py); tests/test_cpu_oracle_and_host.py checks this
restatement against them.

Canonical result (shared with minsu3d_b200.postprocess): dict of
  proposal [n] int32 (original proposal id of each instance, in the reference's output order), label_id [n] int64,
  conf [n] float32, bbox [n,6] float32, mask_points [sum] int32 (ascending inside an instance), mask_offsets [n+1].
The RLE string of the reference (`rle_encode`) is a function of mask_points and is not part of the hot path.
Ties in the NMS score order: the reference uses numpy's unstable argsort; here ties keep the lower proposal first.
"""
import numpy as np


def _sigmoid(x):
    x = np.asarray(x, np.float32).reshape(-1)
    return (np.float32(1) / (np.float32(1) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def nms(cross_ious, scores, threshold):
    """pointgroup.py:197-218."""
    ixs = np.argsort(-scores, kind="stable")
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        ious = cross_ious[i, ixs[1:]]
        remove = np.where(ious > threshold)[0] + 1
        ixs = np.delete(ixs, remove)
        ixs = np.delete(ixs, 0)
    return np.array(pick, dtype=np.int32)


def _instances(masks, proposal_ids, conf, xyz, sem_labels, num_ignored):
    n = masks.shape[0]
    out = {"proposal": np.asarray(proposal_ids, np.int32), "label_id": np.zeros(n, np.int64),
           "conf": np.asarray(conf, np.float32), "bbox": np.zeros((n, 6), np.float32)}
    pts, offs = [], [0]
    for i in range(n):
        m = masks[i]
        out["label_id"][i] = int(sem_labels[m][0]) - num_ignored + 1
        p = xyz[m]
        out["bbox"][i] = np.concatenate((p.min(0), p.max(0)))
        pts.append(np.nonzero(m)[0])
        offs.append(offs[-1] + pts[-1].size)
    out["mask_points"] = (np.concatenate(pts) if pts else np.zeros(0)).astype(np.int32)
    out["mask_offsets"] = np.asarray(offs, np.int32)
    return out


def pointgroup_pred_instances(xyz, proposals_scores, proposals_idx, num_proposals, sem_labels, num_ignored,
                              score_thr, npoint_thr, nms_thr):
    """pointgroup.py:220-265 (sem_labels = semantic_scores.max(1)[1])."""
    score = _sigmoid(proposals_scores)
    n = sem_labels.shape[0]
    mask = np.zeros((num_proposals, n), bool)
    mask[proposals_idx[:, 0], proposals_idx[:, 1]] = True
    npoint = mask.sum(1)
    keep = (score > np.float32(score_thr)) & (npoint > npoint_thr)
    ids = np.nonzero(keep)[0]
    score, mask = score[keep], mask[keep]
    if score.shape[0] == 0:
        pick = np.zeros(0, np.int64)
    else:
        mf = mask.astype(np.float32)
        inter = mf @ mf.T
        npf = mf.sum(1)
        cross = inter / (npf[:, None] + npf[None, :] - inter)
        pick = nms(cross, score, np.float32(nms_thr))
    return _instances(mask[pick], ids[pick], score[pick], xyz, sem_labels, num_ignored)


def hais_pred_instances(xyz, scores, proposals_idx, num_proposals, mask_scores, sem_labels, num_ignored,
                        mask_thr, score_thr, npoint_thr):
    """hais.py:210-247."""
    score = _sigmoid(scores)
    n = sem_labels.shape[0]
    mask = np.zeros((num_proposals, n), bool)
    ok = np.asarray(mask_scores, np.float32).reshape(-1) > np.float32(mask_thr)
    mask[proposals_idx[ok][:, 0], proposals_idx[ok][:, 1]] = True
    ids = np.arange(num_proposals)
    sel = score > np.float32(score_thr)
    score, mask, ids = score[sel], mask[sel], ids[sel]
    sel = mask.sum(1) >= npoint_thr
    score, mask, ids = score[sel], mask[sel], ids[sel]
    return _instances(mask, ids, score, xyz, sem_labels, num_ignored)


def _softmax(x):
    x = np.asarray(x, np.float32)
    e = np.exp(x - x.max(1, keepdims=True), dtype=np.float32)
    return (e / e.sum(1, keepdims=True, dtype=np.float32)).astype(np.float32)


def _concat(parts, n_label_dtype=np.int64):
    out = {"proposal": np.concatenate([p["proposal"] for p in parts]) if parts else np.zeros(0, np.int32),
           "label_id": np.concatenate([p["label_id"] for p in parts]) if parts else np.zeros(0, n_label_dtype),
           "conf": np.concatenate([p["conf"] for p in parts]) if parts else np.zeros(0, np.float32),
           "bbox": np.concatenate([p["bbox"] for p in parts]) if parts else np.zeros((0, 6), np.float32),
           "mask_points": np.concatenate([p["mask_points"] for p in parts]) if parts else np.zeros(0, np.int32)}
    offs = [np.zeros(1, np.int64)]
    base = 0
    for p in parts:
        offs.append(p["mask_offsets"][1:].astype(np.int64) + base)
        base += int(p["mask_offsets"][-1])
    out["mask_offsets"] = np.concatenate(offs).astype(np.int32)
    return out


def softgroup_pred_instances(xyz, proposals_idx, num_points, cls_scores, iou_scores, mask_scores, instance_classes,
                             mask_thr, cls_thr, min_npoint):
    """softgroup.py:269-313: per instance class, proposals above the class-score threshold with at least
    min_npoint points whose class mask score passes; conf = softmax class score * clamp(iou score, 0, 1);
    label_id = class + 1; output ordered by class, then proposal."""
    num_instances = cls_scores.shape[0]
    cls = _softmax(cls_scores)
    parts = []
    for i in range(instance_classes):
        score = cls[:, i] * np.clip(np.asarray(iou_scores, np.float32)[:, i], 0, 1)
        mask = np.zeros((num_instances, num_points), bool)
        ok = np.asarray(mask_scores, np.float32)[:, i] > np.float32(mask_thr)
        mask[proposals_idx[ok][:, 0], proposals_idx[ok][:, 1]] = True
        ids = np.arange(num_instances)
        sel = cls[:, i] > np.float32(cls_thr)
        score, mask, ids = score[sel], mask[sel], ids[sel]
        sel = mask.sum(1) >= min_npoint
        score, mask, ids = score[sel], mask[sel], ids[sel]
        part = _instances(mask, ids, score, xyz, np.zeros(num_points, np.int64), 0)
        part["label_id"][:] = i + 1
        parts.append(part)
    return _concat(parts)
