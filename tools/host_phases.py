"""Host-side time per phase of a PointGroup train step (wall clock, no synchronisation inside the phases)."""
import sys, time; sys.path.insert(0, '.')
import torch
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev)
pool = [scenes.make_batch([4*i, 4*i+1, 4*i+2, 4*i+3], dev, 100_000) for i in range(3)]
for i in range(9): tr.step(pool[i % 3])
torch.cuda.synchronize()
import inspect
print(inspect.getsource(train.Trainer.step))
m, opt = tr.model, tr.optimizer
acc = {}
def tick(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
N = 9
fw = []
import gc, os
if os.environ.get('B2S_NOGC'):
    gc.collect(); gc.freeze(); gc.disable()
for i in range(N):
    data = pool[i % 3]
    torch.cuda.synchronize()
    t = time.perf_counter()
    t00 = t; out = m(data); t = tick("forward (backbone + clustering + scorenet)", t); fw.append(round((t - t00) * 1e3, 1))
    losses = m.loss(data, out); loss = sum(losses.values()); t = tick("loss", t)
    opt.zero_grad(set_to_none=True); loss.backward(); t = tick("backward", t)
    opt.step(); t = tick("optimizer", t)
    torch.cuda.synchronize(); t = tick("drain", t)
print("forward per step:", fw, "reserved GB %.2f" % (torch.cuda.memory_reserved() / 2**30))
for k, v in acc.items(): print("%-45s %6.2f ms" % (k, v / N * 1e3))
