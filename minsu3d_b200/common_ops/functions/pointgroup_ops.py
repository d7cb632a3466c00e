"""Mirror of minsu3d/common_ops/functions/pointgroup_ops.py: pg_bfs_cluster on the GPU."""
import torch

from ... import ops


def pg_bfs_cluster(semantic_label, ball_query_idxs, start_len, threshold):
    """(cluster_idxs [sumNPoint,2] i32, cluster_offsets [nCluster+1] i32), BFS visit order.

    Accepts CUDA tensors (fast path, results stay on the device) or the CPU tensors the reference
    model code passes (pointgroup.py:49-52); in that case results come back on the CPU.
    """
    with torch.no_grad():
        on_cpu = not start_len.is_cuda
        lab, nb, sl = (t.cuda() if not t.is_cuda else t for t in (semantic_label, ball_query_idxs, start_len))
        comp = ops.cluster_label(nb, sl, lab)
        ci, co = ops.cluster_extract(nb, sl, lab, comp, mode=0, thr_i=int(threshold))
        return (ci.cpu(), co.cpu()) if on_cpu else (ci, co)
