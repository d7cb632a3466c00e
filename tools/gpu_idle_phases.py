"""GPU busy / idle time per phase of one PointGroup train step (torch profiler trace, union of kernel intervals over all
streams; phases are delimited by landmark kernels).  The profiler slows the host: read the SHARES."""
import sys, json, os; sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev)
pool = [scenes.make_batch([4*i, 4*i+1, 4*i+2, 4*i+3], dev, 100_000) for i in range(3)]
for i in range(9): tr.step(pool[i % 3])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3): tr.step(pool[i])
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/_idle_trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
os.remove(path)
ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
# steps are delimited by the fused Adam kernel
adam = [i for i, e in enumerate(ks) if "FusedOptimizer" in e["name"] or "fused_adam" in e["name"].lower()]
ends = []
for i in adam:
    if not ends or i - ends[-1] > 50: ends.append(i)
    else: ends[-1] = i
assert len(ends) >= 3, len(ends)
step = ks[ends[0] + 1: ends[1] + 1]  # the middle step
marks = [("backbone forward", None), ("ball query", "bq_keys_kernel"), ("clustering", "cc_init_kernel"), ("cluster voxelise + ScoreNet", "cv_cluster_kernel"),
         ("losses", "ce_forward_kernel"), ("backward", "ce_backward_kernel"), ("optimizer", "FusedOptimizer")]
idx = [0]
for name, key in marks[1:]:
    j = next((i for i, e in enumerate(step) if key in e["name"] and i >= idx[-1]), None)
    idx.append(j if j is not None else idx[-1])
idx.append(len(step))
t_begin = ks[ends[0]]["ts"] + ks[ends[0]]["dur"]
print("%-30s %8s %8s %8s %8s" % ("phase", "span ms", "busy ms", "idle ms", "kernels"))
tot = [0, 0, 0]
for p, (name, _) in enumerate(marks):
    seg = step[idx[p]: idx[p + 1]]
    if not seg: continue
    t0 = t_begin if p == 0 else step[idx[p]]["ts"]
    t1 = step[idx[p + 1]]["ts"] if idx[p + 1] < len(step) else max(e["ts"] + e["dur"] for e in seg)
    busy, cur = 0.0, t0
    for e in seg:
        a, b = max(e["ts"], cur), min(e["ts"] + e["dur"], t1)
        if b > a: busy += b - a; cur = b
    print("%-30s %8.2f %8.2f %8.2f %8d" % (name, (t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(seg)))
    tot[0] += t1 - t0; tot[1] += busy; tot[2] += len(seg)
print("%-30s %8.2f %8.2f %8.2f %8d" % ("step", tot[0] / 1e3, tot[1] / 1e3, (tot[0] - tot[1]) / 1e3, tot[2]))
