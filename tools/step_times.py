"""Per-step wall time (synchronised) of the PointGroup train step: reveals outliers."""
import sys, time, os; sys.path.insert(0, os.getcwd())
import torch
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev, reserve_gb=float(os.environ.get('B2S_RESERVE', '0')))
pool = [scenes.make_batch([4*i, 4*i+1, 4*i+2, 4*i+3], dev, 100_000) for i in range(3)]
for i in range(9): tr.step(pool[i % 3])
torch.cuda.synchronize()
ts = []
for i in range(40):
    t = time.perf_counter(); tr.step(pool[i % 3]); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
print("median %.1f  mean %.1f  max %.1f  slow(>1.5x median): %s" % (sorted(ts)[20], sum(ts) / 40, max(ts),
      [(i, round(t)) for i, t in enumerate(ts) if t > 1.5 * sorted(ts)[20]]))
