"""Mirror of minsu3d/common_ops/functions/softgroup_ops.py: sg_bfs_cluster, global_avg_pool."""
import torch
from torch.autograd import Function

from ... import ops


def sg_bfs_cluster(class_numpoint_mean, ball_query_idxs, start_len, threshold, class_id):
    with torch.no_grad():
        on_cpu = not start_len.is_cuda
        nb, sl = (t.cuda() if not t.is_cuda else t for t in (ball_query_idxs, start_len))
        mean = torch.tensor(class_numpoint_mean, dtype=torch.float32)[class_id]
        thr = torch.tensor(threshold, dtype=torch.float32)
        if float(mean) != -1:  # bfs_cluster.cpp:116-121
            thr = thr * mean
        comp = ops.cluster_label(nb, sl, None)
        ci, co = ops.cluster_extract(nb, sl, None, comp, mode=1, thr_f=float(thr))
        return (ci.cpu(), co.cpu()) if on_cpu else (ci, co)


class GlobalAvgPool(Function):
    @staticmethod
    def forward(ctx, feats, proposals_offset):
        n_prop = proposals_offset.size(0) - 1
        sum_npoint, c = feats.size()
        out = torch.empty((n_prop, c), dtype=torch.float32, device=feats.device)
        ops.sec_reduce("avg", feats, proposals_offset, out)
        ctx.for_backwards = (proposals_offset, sum_npoint)
        return out

    @staticmethod
    def backward(ctx, d_output_feats):
        proposals_offset, sum_npoint = ctx.for_backwards
        c = d_output_feats.size(1)
        d_feats = torch.zeros((sum_npoint, c), dtype=torch.float32, device=d_output_feats.device)
        ops.global_avg_pool_bp(d_feats, proposals_offset, d_output_feats.contiguous())
        return d_feats, None


global_avg_pool = GlobalAvgPool.apply
