import sys, time, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops
import oracle
rng = np.random.default_rng(0)
c = surface_voxels(rng, 330000, batch=4); n = c.shape[0]
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
table, _, _, oc = ops.coord_unique(D(c), 1); nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
if os.environ.get('B2S_NOMASK'): tmask = None
if os.environ.get('B2S_LOCAL'):
    m = int(os.environ['B2S_LOCAL'])
    nbr = torch.where(nbr >= 0, nbr % m, nbr).contiguous()  # latency probe: every gathered row is L1/L2-hot
print('rows', n, 'rho', float((nbr >= 0).sum()) / n, 'flags', os.environ.get('B2S_TC_DEBUG'))
for (cin, cout) in [(16,16),(32,32),(64,64)]:
    x = torch.randn(n, cin, device='cuda'); w = torch.randn(27, cin, cout, device='cuda') * 0.05
    ref = ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=1)
    for algo in (2, 3):
        for _ in range(3): y = ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=algo, tile_mask=tmask)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=algo, tile_mask=tmask)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
        err = float((y - ref).abs().max() / ref.abs().max())
        print(cin, cout, 'algo', algo, '%.1f us' % (dt * 1e6), 'err %.2e' % err, flush=True)
