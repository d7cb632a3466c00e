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
    import time
    t = time.perf_counter(); 
    for _ in range(3): tr.step(data)
    torch.cuda.synchronize(); print("wall ms/step", (time.perf_counter() - t) / 3 * 1e3)
