import sys, time, cProfile, pstats; sys.path.insert(0, '.')
import torch
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev)
pool = [scenes.make_batch([4*i, 4*i+1, 4*i+2, 4*i+3], dev, 100_000) for i in range(3)]
for i in range(9): tr.step(pool[i % 3])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(6): tr.step(pool[i % 3])
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(45)
