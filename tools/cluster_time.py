"""Per-call timing of the clustering stage inside one PointGroup train step (synchronised wall clock)."""
import sys, time; sys.path.insert(0, '.')
import torch
from minsu3d_b200 import ops
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev)
data = scenes.make_batch([0, 1, 2, 3], dev, 100_000)
for _ in range(3): tr.step(data)
torch.cuda.synchronize()
log = []
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); log.append((name, (time.perf_counter() - t) * 1e3, r))
        return r
    setattr(mod, name, g)
for nm in ("ballquery", "cluster_label", "cluster_extract"):
    wrap(ops, nm)
tr.step(data); torch.cuda.synchronize()
for name, ms, r in log:
    extra = ""
    if name == "ballquery": extra = "n=%d nActive=%d max_len=%d" % (r[1].size(0), r[0].numel(), int(r[1][:, 1].max()))
    if name == "cluster_extract": extra = "clusters=%d points=%d" % (r[1].numel() - 1, r[0].size(0))
    print("%-16s %7.3f ms  %s" % (name, ms, extra))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(data); torch.cuda.synchronize()
print("-- clustering kernels in one step (us total / calls)")
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if any(t in e.key for t in ("bq_", "cc_", "cl_", "bfs_", "cluster", "DeviceRadixSort", "DeviceScan", "ha_")):
        print("%-60s %9.1f %4d" % (e.key[:60], e.device_time_total, e.count))
