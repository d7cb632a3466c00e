"""Sorted-tile schedule vs row order on the benchmark batch's level-0 map (run on the GPU box)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from minsu3d_b200 import ops
from minsu3d_b200.harness import scenes
d = scenes.make_batch([0, 1, 2, 3], torch.device('cuda'))
table, _, _, oc = ops.coord_unique(d['voxel_xyz'], 1); n = oc.size(0)
nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
perm, nbs, tms = ops.tile_order(nbr)
popc = lambda tm: sum(bin(v & 0xFFFFFFFF).count('1') for v in tm.tolist()) / tm.numel()
print('rows', n, 'rho %.2f' % (float((nbr >= 0).sum()) / n), 'active/tile row-order %.2f sorted %.2f' % (popc(tmask), popc(tms)))
def bench(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
print('tile_order %.1f us' % bench(lambda: ops.tile_order(nbr)))
for (cin, cout) in [(16, 16), (32, 32), (64, 64)]:
    x = torch.randn(n, cin, device='cuda'); w = torch.randn(27, cin, cout, device='cuda') * 0.05
    for algo in (2, 3):
        a = bench(lambda: ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=algo, tile_mask=tmask))
        b = bench(lambda: ops.conv_table(x, w, nbs, n, 27, cin, cout, algo=algo, tile_mask=tms, out_rows=perm))
        print(cin, cout, 'algo', algo, 'row-order %.1f us  sorted %.1f us' % (a, b), flush=True)
