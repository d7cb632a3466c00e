"""Drop-in for the reference's `COMMON_OPS` pybind module (boundary #2).

Same 15 names and positional argument orders as
minsu3d/common_ops/src/common_ops_api.cpp:6-29, so the reference's own
minsu3d/common_ops/functions/*.py run unmodified on top of it after
`minsu3d_b200.install_as_reference_modules()` (sys.modules["COMMON_OPS"] = this module).

Every function computes on the GPU through libb2s.  The reference's BFS / hierarchical
aggregation take CPU tensors (the model code calls `.cpu()` first, pointgroup.py:41-52): such
inputs are uploaded, processed on the GPU and the caller's CPU output tensors are resized and
filled, exactly like `resize_()` in bfs_cluster.cpp:157-158.  CUDA inputs stay on the device.
"""
import torch

from . import ops

I32 = torch.int32


def _cuda(t):
    return t if t.is_cuda else t.cuda()


def _fill_like_resize(dst, src):
    """dst.resize_(src.shape); dst.copy_(src)  (reference: tensor.resize_ + fill in place)."""
    dst.resize_(src.shape)
    dst.copy_(src)


# ---- ball query (bfs_cluster.cpp:15-25) ------------------------------------------------------
def ballquery_batch_p(xyz, batch_idxs, batch_offsets, idx, start_len, n, meanActive, radius):
    found, sl = ops.ballquery(xyz, batch_idxs, batch_offsets, radius)
    start_len.copy_(sl)
    n_active = found.numel()
    if n_active <= idx.numel():
        idx[:n_active].copy_(found)
    return n_active  # > n*meanActive makes the reference wrapper retry (common_ops.py:31-37)


# ---- BFS clustering (bfs_cluster.cpp:147-187) ------------------------------------------------
def pg_bfs_cluster(semantic_label, ball_query_idxs, start_len, cluster_idxs, cluster_offsets, N, threshold):
    lab, nb, sl = _cuda(semantic_label), _cuda(ball_query_idxs), _cuda(start_len)
    comp = ops.cluster_label(nb, sl, lab)
    ci, co = ops.cluster_extract(nb, sl, lab, comp, mode=0, thr_i=int(threshold))
    _fill_like_resize(cluster_idxs, ci)
    _fill_like_resize(cluster_offsets, co)


def sg_bfs_cluster(class_numpoint_mean, ball_query_idxs, start_len, cluster_idxs, cluster_offsets, N, threshold,
                   class_id):
    nb, sl = _cuda(ball_query_idxs), _cuda(start_len)
    mean = float(class_numpoint_mean[class_id])
    # bfs_cluster.cpp:116-121: thr = threshold if mean == -1 else threshold * mean (fp32 product)
    thr = torch.tensor(threshold, dtype=torch.float32)
    if mean != -1:
        thr = thr * torch.tensor(mean, dtype=torch.float32)
    comp = ops.cluster_label(nb, sl, None)
    ci, co = ops.cluster_extract(nb, sl, None, comp, mode=1, thr_f=float(thr))
    _fill_like_resize(cluster_idxs, ci)
    _fill_like_resize(cluster_offsets, co)


# ---- hierarchical aggregation (hierarchical_aggregation.cpp:108-183) ---------------------------
def hierarchical_aggregation(semantic_label, coord_shift, batch_idxs, ball_query_idxs, start_len,
                             fragment_idxs, fragment_offsets, fragment_centers,
                             cluster_idxs_kept, cluster_offsets_kept, cluster_centers_kept,
                             primary_idxs, primary_offsets, primary_centers,
                             primary_idxs_post, primary_offsets_post,
                             point_num_avg, radius_avg, N, using_set_aggr_, ignored_label):
    lab, xyz, bidx = _cuda(semantic_label), _cuda(coord_shift), _cuda(batch_idxs)
    nb, sl = _cuda(ball_query_idxs), _cuda(start_len)
    pna, rad = _cuda(point_num_avg), _cuda(radius_avg)
    comp = ops.cluster_label(nb, sl, lab)

    def group(g):
        ci, co = ops.cluster_extract(nb, sl, lab, comp, mode=2, point_num_avg=pna, group=g)
        cc = ops.cluster_centers(ci, co, xyz, lab, bidx)
        return ci, co, cc

    k_i, k_o, k_c = group(1)
    p_i, p_o, p_c = group(2)
    for dst, src in ((cluster_idxs_kept, k_i), (cluster_offsets_kept, k_o), (cluster_centers_kept, k_c),
                     (primary_idxs, p_i), (primary_offsets, p_o), (primary_centers, p_c)):
        _fill_like_resize(dst, src)
    if int(using_set_aggr_) == 0:
        return
    f_i, f_o, f_c = group(3)
    for dst, src in ((fragment_idxs, f_i), (fragment_offsets, f_o), (fragment_centers, f_c)):
        _fill_like_resize(dst, src)
    n_prim = p_o.numel() - 1
    total = f_i.size(0) + p_i.size(0)
    if n_prim == 0:  # hierarchical_aggregation.cu:98-100: early return, outputs stay zero
        primary_idxs_post.resize_((total, 2)).zero_()
        primary_offsets_post.resize_((1,)).zero_()
        return
    post_i, post_o, _ = ops.ha_set_aggregate(f_i, f_o, f_c, p_i, p_o, p_c, rad)
    _fill_like_resize(primary_idxs_post, post_i)   # (sumNPoint_fragment + sumNPoint_primary, 2), tail unused
    _fill_like_resize(primary_offsets_post, post_o)


# ---- segmented ops ---------------------------------------------------------------------------
def sec_mean(inp, offsets, out, nProposal, C):
    ops.sec_reduce("mean", inp, offsets, out)


def sec_min(inp, offsets, out, nProposal, C):
    ops.sec_reduce("min", inp, offsets, out)


def sec_max(inp, offsets, out, nProposal, C):
    ops.sec_reduce("max", inp, offsets, out)


def roipool_fp(feats, proposals_offset, output_feats, output_maxidx, nProposal, C):
    ops.roipool_fp(feats, proposals_offset, output_feats, output_maxidx)


def roipool_bp(d_feats, proposals_offset, output_maxidx, d_output_feats, nProposal, C):
    ops.roipool_bp(d_feats, proposals_offset, output_maxidx, d_output_feats)


def global_avg_pool_fp(feats, proposals_offset, output_feats, nProposal, C):
    ops.sec_reduce("avg", feats, proposals_offset, output_feats)


def global_avg_pool_bp(d_feats, proposals_offset, d_output_feats, nProposal, C):
    ops.global_avg_pool_bp(d_feats, proposals_offset, d_output_feats)


# ---- IoU / mask labels -----------------------------------------------------------------------
def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, nInstance, nProposal):
    ops.get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou)


def get_mask_iou_on_cluster(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou,
                            nInstance, nProposal):
    ops.get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou)


def get_mask_iou_on_pred(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou,
                         nInstance, nProposal, mask_scores_sigmoid):
    ops.get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou,
                mask_scores=mask_scores_sigmoid.reshape(-1))


def get_mask_label(proposals_idx, proposals_offset, instance_labels, instance_cls, proposals_iou, nInstance,
                   nProposal, ignored_label, iou_thr, mask_label, mask_label_mask):
    ops.get_mask_label(proposals_idx, proposals_offset, instance_labels, instance_cls, proposals_iou, nInstance,
                       nProposal, ignored_label, iou_thr, mask_label, mask_label_mask)
