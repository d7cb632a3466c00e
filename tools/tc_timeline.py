import sys, os, ctypes; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ['B2S_TC_DEBUG'] = sys.argv[3] if len(sys.argv) > 3 else '16'
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops, _cabi
rng = np.random.default_rng(0)
c = surface_voxels(rng, 330000, batch=4); n = c.shape[0]
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
table, _, _, oc = ops.coord_unique(D(c), 1); nbr = ops.kernel_map(oc, table, 3, 1)
cin = cout = int(sys.argv[1]); algo = int(sys.argv[2])
x = torch.randn(n, cin, device='cuda'); w = torch.randn(27, cin, cout, device='cuda') * 0.05
for _ in range(3): y = ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=algo)
torch.cuda.synchronize()
buf = np.zeros((4, 256), np.int64)
_cabi.lib().b2s_debug_tc_timeline(buf.ctypes.data)
t0 = buf[0][0]
T = int((buf[1] > 0).sum())
print('slabs recorded', T)
print('slab  p_wait_done  p_arrive  mma_wake  mma_commit   (cycles since first producer wait)')
for i in range(min(T, 40)):
    print('%4d %11d %9d %9d %10d' % (i, buf[0][i]-t0, buf[1][i]-t0, buf[2][i]-t0, buf[3][i]-t0))
d = np.diff(buf[1][:T]); print('producer arrive period: mean %.0f median %.0f' % (d.mean(), np.median(d)))
print('mma wake - producer(thread0) arrive: mean %.0f' % (buf[2][:T]-buf[1][:T]).mean())
print('mma commit - wake: mean %.0f' % (buf[3][:T]-buf[2][:T]).mean())
