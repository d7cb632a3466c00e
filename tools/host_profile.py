"""Host-side profile of the PointGroup train step: enqueue vs total time, per-phase host time, cProfile."""
import sys, time; sys.path.insert(0, '.')
import cProfile, pstats, io
import torch
from minsu3d_b200 import _cabi
from minsu3d_b200.harness import models, scenes, train
dev = torch.device("cuda", 0)
cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
tr = train.Trainer(cfg, dev, reserve_gb=16.0)
pool = [scenes.make_batch([4*i, 4*i+1, 4*i+2, 4*i+3], dev, 100_000) for i in range(3)]
for i in range(9): tr.step(pool[i % 3])
torch.cuda.synchronize()
enq, tot = [], []
for i in range(9):
    t0 = time.perf_counter(); tr.step(pool[i % 3]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    enq.append(t1 - t0); tot.append(t2 - t0)
print('enqueue ms', [round(x*1e3,1) for x in enq]); print('total ms', [round(x*1e3,1) for x in tot])
_cabi.reset_launch_count(); tr.step(pool[0]); torch.cuda.synchronize(); print('libb2s kernels per step', _cabi.launch_count())
m, opt = tr.model, tr.optimizer
acc = {}
def tick(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
N = 9
for i in range(N):
    data = pool[i % 3]
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = m(data); t = tick("forward", t)
    losses = m.loss(data, out); loss = sum(losses.values()); t = tick("loss", t)
    opt.zero_grad(set_to_none=True); loss.backward(); t = tick("backward", t)
    opt.step(); t = tick("optimizer", t)
    torch.cuda.synchronize(); t = tick("drain", t)
for k, v in acc.items(): print("%-12s %6.2f ms" % (k, v / N * 1e3))
pr = cProfile.Profile(); pr.enable()
for i in range(3): tr.step(pool[i % 3])
torch.cuda.synchronize(); pr.disable()
for key, n in (('tottime', 45), ('cumulative', 70)):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(n); print(s.getvalue())
