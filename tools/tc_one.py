import sys, time, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops
rng = np.random.default_rng(0)
c = surface_voxels(rng, 330000, batch=4); n = c.shape[0]
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
table, _, _, oc = ops.coord_unique(D(c), 1); nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
if os.environ.get('B2S_NOMASK'): tmask = None
cin = cout = int(sys.argv[1]); algo = int(sys.argv[2])
x = torch.randn(n, cin, device='cuda'); w = torch.randn(27, cin, cout, device='cuda') * 0.05
for _ in range(3): y = ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=algo, tile_mask=tmask)
torch.cuda.synchronize()
