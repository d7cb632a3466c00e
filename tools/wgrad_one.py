import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops
rng = np.random.default_rng(0)
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ch = int(sys.argv[1]); rows = int(sys.argv[2])
c = surface_voxels(rng, rows, batch=4); n = c.shape[0]
table, _, _, oc = ops.coord_unique(D(c), 1); nbr = ops.kernel_map(oc, table, 3, 1)
pin, pout, koff, _ = ops.pairs_from_nbr(nbr)
x = torch.randn(n, ch, device='cuda'); g = torch.randn(n, ch, device='cuda')
for _ in range(3): gw = ops.conv_wgrad(x, g, pin, pout, koff, 27, ch, ch, n * 27)
torch.cuda.synchronize()
