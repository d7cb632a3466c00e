"""GPU idle gaps inside one train step: exports a torch profiler trace and lists the largest gaps between
consecutive kernels together with the kernels around them."""
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
    tr.step(pool[0]); tr.step(pool[1])
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/step_trace.json")
ev = json.load(open("gpurun_out/step_trace.json"))["traceEvents"]
ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0, t1 = ks[0]["ts"], ks[-1]["ts"] + ks[-1]["dur"]
busy = sum(e["dur"] for e in ks)
print("span %.2f ms, busy %.2f ms, kernels %d" % ((t1 - t0) / 1e3, busy / 1e3, len(ks)))
gaps = []
for a, b in zip(ks, ks[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 0: gaps.append((g, a["name"][:50], b["name"][:50], (a["ts"] - t0) / 1e3))
tot = sum(g[0] for g in gaps)
print("total gap %.2f ms; gaps > 20us: %.2f ms (%d); gaps <= 20us: %.2f ms" % (
    tot / 1e3, sum(g[0] for g in gaps if g[0] > 20) / 1e3, sum(1 for g in gaps if g[0] > 20),
    sum(g[0] for g in gaps if g[0] <= 20) / 1e3))
for g in sorted(gaps, reverse=True)[:40]:
    print("%8.1f us at %7.2f ms  after %-50s before %s" % (g[0], g[3], g[1], g[2]))
