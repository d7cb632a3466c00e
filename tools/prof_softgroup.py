"""SoftGroup inference (configs[3]): wall time per scene, GPU busy time and the top kernels (torch profiler)."""
import sys, time; sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from minsu3d_b200 import postprocess
from minsu3d_b200.harness import models, scenes
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("softgroup", proposal_source="gt_noise")
torch.manual_seed(123)
model = models.build_model(cfg).to(dev).eval()
sizes = [50_000, 100_000, 150_000, 200_000, 250_000, 120_000, 80_000, 180_000]
distinct = [scenes.collate([scenes.make_scene(500 + i, n)], dev) for i, n in enumerate(sizes)]
inst_classes = cfg.classes - len(cfg.ignore_classes)
def one(d, t=None):
    with torch.no_grad():
        t0 = time.perf_counter()
        out = model(d)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        res = postprocess.softgroup_pred_instances(d["point_xyz"], out["proposals_idx"], d["point_xyz"].size(0), out["cls_scores"],
                                                   out["iou_scores"], out["mask_scores"], inst_classes, -0.5, 0.001, 100)
        n = int(res["label_id"].numel())
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if t is not None: t.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, d["point_xyz"].size(0), int(out["proposals_idx"].size(0)), n))
for d in distinct: one(d)
for d in distinct: one(d)
ts = []
for d in distinct: one(d, ts)
for r in ts: print("model %6.1f ms  postproc %6.1f ms  points %7d  proposal rows %8d  instances %d" % r)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for d in distinct: one(d)
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and not e.key.startswith(("autograd", "aten::", "cuda"))]
print("GPU busy per scene %.2f ms, launches per scene %.0f" % (sum(r[2] for r in rows) / 8e3, sum(r[1] for r in rows) / 8))
for k, c, t in sorted(rows, key=lambda r: -r[2])[:28]:
    print("%7.1f %9.1f us  %s" % (c / 8, t / 8, k[:100]))
