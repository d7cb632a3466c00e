"""ncu target: the map-building and clustering kernels on the benchmark batch (4 x 100k points: 325k voxels, 136k foreground
points): coordinate insert + kernel map of the level-0 map, ball query (count, fill, bitmap fill) and clustering (union-find
hook, directed labels, BFS order) on raw and on shifted coordinates.  One pass inside the profiler range."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from minsu3d_b200 import ops
from minsu3d_b200.harness import models, scenes
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
data = scenes.make_batch([0, 1, 2, 3], dev, 100_000)
model = models.build_model(cfg).to(dev)
n_pts = data["point_xyz"].size(0)
scores, offsets = model._cluster_inputs(data, {"semantic_scores": torch.zeros((n_pts, cfg.classes), device=dev)})
preds = scores.argmax(1).to(torch.int16)
obj = model._object_points(preds)
bidx = data["vert_batch_ids"][obj].contiguous()
boffs = torch.cumsum(torch.bincount(bidx + 1), dim=0).int()
lab = preds[obj].contiguous()
sets = [data["point_xyz"][obj].contiguous(), (data["point_xyz"] + offsets)[obj].contiguous()]

def run():
    table, _, _, oc = ops.coord_unique(data["voxel_xyz"], 1)
    ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
    for pts in sets:
        idx, sl = ops.ballquery(pts, bidx, boffs, cfg.cluster_radius)
        comp = ops.cluster_label(idx, sl, lab)
        ops.cluster_extract(idx, sl, lab, comp, mode=0, thr_i=cfg.cluster_npoint_thre)

run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
