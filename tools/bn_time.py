import sys, time; sys.path.insert(0, '.')
import torch
from minsu3d_b200 import ops
dev = 'cuda'
def bench(f, iters=20):
    # device time per call: CUDA events around a batch of launches (launches are queued faster than they run for
    # the large shapes; for the small ones this reports the launch-bound rate)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / iters * 1e3
for n, c in [(330000, 16), (330000, 32), (90000, 32), (90000, 64), (25000, 48), (25000, 96), (7000, 64), (2000, 80), (500, 112)]:
    x = torch.randn(n, c, device=dev); dy = torch.randn(n, c, device=dev)
    g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev); rm = torch.zeros(c, device=dev); rv = torch.ones(c, device=dev)
    t_stats = bench(lambda: ops.bn_stats(x, 1e-4, 0.1, rm, rv))
    mean, rstd = ops.bn_stats(x, 1e-4, 0.1, rm, rv)
    y = ops.bn_apply(x, mean, rstd, g, b, True)
    t_apply = bench(lambda: ops.bn_apply(x, mean, rstd, g, b, True))
    t_bwd = bench(lambda: ops.bn_backward(x, y, dy, mean, rstd, g, True, True))
    mb = n * c * 4 / 1e6
    print("n=%6d c=%3d  stats %6.1f us (%5.0f GB/s)  apply %6.1f us (%5.0f GB/s)  backward %6.1f us (%5.0f GB/s)" % (
        n, c, t_stats, mb / t_stats * 1e3, t_apply, 2 * mb / t_apply * 1e3, t_bwd, 7 * mb / t_bwd * 1e3))
from torch.profiler import profile, ProfilerActivity
print("-- kernel durations (torch profiler, us)")
for n, c in [(330000, 16), (330000, 32), (90000, 64), (7000, 64)]:
    x = torch.randn(n, c, device=dev); dy = torch.randn(n, c, device=dev)
    g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev); rm = torch.zeros(c, device=dev); rv = torch.ones(c, device=dev)
    mean, rstd = ops.bn_stats(x, 1e-4, 0.1, rm, rv); y = ops.bn_apply(x, mean, rstd, g, b, True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.bn_stats(x, 1e-4, 0.1, rm, rv); ops.bn_apply(x, mean, rstd, g, b, True); ops.bn_backward(x, y, dy, mean, rstd, g, True, True)
        torch.cuda.synchronize()
    print(n, c, {e.key.split('(')[0][-24:]: round(e.device_time_total / e.count, 1) for e in prof.key_averages() if 'bn_' in e.key})
