import sys, time; sys.path.insert(0, '.')
import torch
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev)
pool = [scenes.make_batch([4*i, 4*i+1, 4*i+2, 4*i+3], dev, 100_000) for i in range(3)]
for i in range(9): tr.step(pool[i % 3])
torch.cuda.synchronize()
enq, tot = [], []
for i in range(9):
    t0 = time.perf_counter(); tr.step(pool[i % 3]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    enq.append(t1 - t0); tot.append(t2 - t0)
print('enqueue ms', [round(x*1e3,1) for x in enq]); print('total ms', [round(x*1e3,1) for x in tot])
# phases: forward-only timing, with syncs
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(3): tr.step(pool[i % 3])
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
