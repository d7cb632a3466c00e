"""Mirror of minsu3d/common_ops/functions/hais_ops.py: hierarchical_aggregation on the GPU."""
import torch

from ... import ops


def hierarchical_aggregation(semantic_label, coord_shift, ball_query_idxs, start_len, batch_idxs, using_set_aggr,
                             point_num_avg, radius_avg, ignored_label):
    """Returns (cluster_idxs [S,2] i32, cluster_offsets [nC+1] i32) = cat(kept fragments, primaries)."""
    with torch.no_grad():
        on_cpu = not start_len.is_cuda
        lab, xyz, nb, sl, bidx = (t.cuda() if not t.is_cuda else t
                                  for t in (semantic_label, coord_shift, ball_query_idxs, start_len, batch_idxs))
        dev = sl.device
        pna = torch.tensor(point_num_avg, dtype=torch.float32, device=dev)
        rad = torch.tensor(radius_avg, dtype=torch.float32, device=dev)
        comp = ops.cluster_label(nb, sl, lab)
        kept_i, kept_o = ops.cluster_extract(nb, sl, lab, comp, mode=2, point_num_avg=pna, group=1)
        prim_i, prim_o = ops.cluster_extract(nb, sl, lab, comp, mode=2, point_num_avg=pna, group=2)
        if int(using_set_aggr) != 0 and prim_o.numel() > 1:
            frag_i, frag_o = ops.cluster_extract(nb, sl, lab, comp, mode=2, point_num_avg=pna, group=3)
            frag_c = ops.cluster_centers(frag_i, frag_o, xyz, lab, bidx)
            prim_c = ops.cluster_centers(prim_i, prim_o, xyz, lab, bidx)
            post_i, post_o, _ = ops.ha_set_aggregate(frag_i, frag_o, frag_c, prim_i, prim_o, prim_c, rad)
            prim_i, prim_o = post_i[:int(post_o[-1])], post_o
        cluster_idxs, cluster_offsets = kept_i, kept_o
        if prim_i.shape[0] != 0:  # hais_ops.py:66-71
            prim_i = prim_i.clone()
            prim_i[:, 0] += cluster_offsets.size(0) - 1
            cluster_idxs = torch.cat((cluster_idxs, prim_i), dim=0)
            cluster_offsets = torch.cat((cluster_offsets, prim_o[1:] + cluster_offsets[-1]))
        return (cluster_idxs.cpu(), cluster_offsets.cpu()) if on_cpu else (cluster_idxs, cluster_offsets)
