import sys, os, time; sys.path.insert(0, os.getcwd())
import torch
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev, reserve_gb=float(os.environ.get('B2S_RESERVE', '0')))
pool = [scenes.make_batch([4*i, 4*i+1, 4*i+2, 4*i+3], dev, 100_000) for i in range(3)]
def st():
    s = torch.cuda.memory_stats()
    return (s.get("num_device_alloc", 0), s.get("num_device_free", 0), s.get("num_alloc_retries", 0),
            round(s["reserved_bytes.all.current"] / 2**20), round(s["active_bytes.all.peak"] / 2**20), round(s["reserved_bytes.all.peak"] / 2**20))
print("conf", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"))
prev = st()
for i in range(36):
    t = time.perf_counter(); tr.step(pool[i % 3]); torch.cuda.synchronize(); d = (time.perf_counter() - t) * 1e3
    cur = st()
    if cur[:2] != prev[:2] or d > 55: print("step %2d %6.1f ms  device_alloc %d device_free %d retries %d reserved %d MB (peak active %d, peak reserved %d)" % ((i, d) + cur))
    prev = cur
