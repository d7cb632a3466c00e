import sys, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops
from torch.profiler import profile, ProfilerActivity
rng = np.random.default_rng(0)
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for rows, cs in ((330000, (16, 32)), (90000, (32, 48)), (25000, (48, 64, 96)), (3000, (80, 112))):
    c = surface_voxels(rng, rows, batch=4); n = c.shape[0]
    table, _, _, oc = ops.coord_unique(D(c), 1); nbr = ops.kernel_map(oc, table, 3, 1)
    pin, pout, koff, _ = ops.pairs_from_nbr(nbr)
    for ch in cs:
        x = torch.randn(n, ch, device='cuda'); g = torch.randn(n, ch, device='cuda')
        ref = None
        for _ in range(2): gw = ops.conv_wgrad(x, g, pin, pout, koff, 27, ch, ch, n * 27)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5): ops.conv_wgrad(x, g, pin, pout, koff, 27, ch, ch, n * 27)
            torch.cuda.synchronize()
        t = [e.device_time_total / e.count for e in prof.key_averages() if 'wgrad' in e.key][0]
        # check against a dense reference for one offset
        k = 13; sel = nbr[:, k] >= 0
        want = x[nbr[sel, k].long()].double().T @ g[sel].double()
        err = float((gw[k].double() - want).abs().max() / want.abs().max())
        print("rows %6d C %3d  wgrad %7.1f us  rel err %.1e" % (n, ch, t, err))
