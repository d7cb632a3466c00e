import sys; sys.path.insert(0, '.')
import torch
from minsu3d_b200 import ops
from torch.profiler import profile, ProfilerActivity
dev = 'cuda'
for n, c in [(330000, 16), (330000, 32), (90000, 32), (90000, 64), (25000, 96), (7000, 64), (500, 112)]:
    x = torch.randn(n, c, device=dev); dy = torch.randn(n, c, device=dev)
    g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev); rm = torch.zeros(c, device=dev); rv = torch.ones(c, device=dev)
    mean, rstd = ops.bn_stats(x, 1e-4, 0.1, rm, rv); y = ops.bn_apply(x, mean, rstd, g, b, True)
    torch.cuda.synchronize()
    res = {}
    for name, f in (("stats", lambda: ops.bn_stats(x, 1e-4, 0.1, rm, rv)), ("bwd", lambda: ops.bn_backward(x, y, dy, mean, rstd, g, True, True))):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5): f()
            torch.cuda.synchronize()
        for e in prof.key_averages():
            if 'bn_' in e.key: res[name + ":" + e.key.split('(')[0][-16:]] = round(e.device_time_total / e.count, 1)
    print(n, c, res)
