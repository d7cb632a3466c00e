import sys, os; sys.path.insert(0, '.')
import torch
from minsu3d_b200.harness import models, scenes, train
mode = sys.argv[1] if len(sys.argv) > 1 else "torch"
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev)
data = scenes.make_batch([0, 1, 2, 3], dev, 100_000)
for _ in range(3):
    tr.step(data)
torch.cuda.synchronize()
if mode == "ncu":
    torch.cuda.cudart().cudaProfilerStart()
    tr.step(data)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
else:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            tr.step(data)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and not e.key.startswith(("autograd", "aten::", "_", "cuda"))]
    print("=== device activities per step (name, launches, us)")
    for k, c, t in sorted(rows, key=lambda r: -r[1]):
        print("%6.1f %9.1f  %s" % (c / 2, t / 2, k[:110]))
    print("launches per step", sum(r[1] for r in rows) / 2)
    import time
    t = time.perf_counter(); 
    for _ in range(3): tr.step(data)
    torch.cuda.synchronize(); print("wall ms/step", (time.perf_counter() - t) / 3 * 1e3)
