"""Weight-gradient kernels: CUDA time per call (all kernels of the call, torch profiler), error vs an fp64 reference over
every offset, and run-to-run bit equality.  B2S_WGRAD_DET=0 / B2S_WGRAD_MMA=0 select the older kernels."""
import sys, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops
from torch.profiler import profile, ProfilerActivity
rng = np.random.default_rng(0)
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
SHAPES = ((330000, ((16, 16), (32, 16), (32, 32))), (90000, ((32, 32), (64, 32), (48, 48))), (25000, ((48, 48), (64, 64), (96, 96))),
          (5000, ((64, 64), (128, 64))), (1000, ((80, 80), (112, 112), (224, 112))), (80, ((112, 112),)))
if os.environ.get('WG_SMALL'): SHAPES = SHAPES[3:]
for rows, cs in SHAPES:
    c = surface_voxels(rng, rows, batch=4); n = c.shape[0]
    table, _, _, oc = ops.coord_unique(D(c), 1); nbr = ops.kernel_map(oc, table, 3, 1)
    pin, pout, koff, _ = ops.pairs_from_nbr(nbr)
    for ca, cg in cs:
        x = torch.randn(n, ca, device='cuda'); g = torch.randn(n, cg, device='cuda')
        for _ in range(2): gw = ops.conv_wgrad(x, g, pin, pout, koff, 27, ca, cg, n * 27)
        gw = gw.clone()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5): gw2 = ops.conv_wgrad(x, g, pin, pout, koff, 27, ca, cg, n * 27)
            torch.cuda.synchronize()
        t = sum(e.device_time_total for e in prof.key_averages() if 'wgrad' in e.key or 'Memset' in e.key) / 5
        same = bool(torch.equal(gw, gw2))
        if os.environ.get("WG_DETAIL"):
            for e in prof.key_averages():
                print("      %-60s x%d  %.1f us" % (e.key[:60], e.count, e.device_time_total / max(e.count, 1)))
        err = 0.0
        for k in range(27):
            sel = nbr[:, k] >= 0
            want = x[nbr[sel, k].long()].double().T @ g[sel].double()
            err = max(err, float((gw[k].double() - want).abs().max() / want.abs().max().clamp_min(1e-30)))
        print("rows %6d C %3d x %3d  wgrad %7.1f us  rel err %.1e  reproducible %s" % (n, ca, cg, t, err, same))
