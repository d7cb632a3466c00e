import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import oracle
from helpers import surface_voxels
from minsu3d_b200 import ops
dev = 'cuda'
def D(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def rel(got, want):
    got = got.detach().cpu().double().numpy(); want = np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
rng = np.random.default_rng(0)
print("== table mode")
for (cin, cout, n) in [(16,16,300), (16,16,20000), (32,16,20000), (32,32,20000), (48,48,8000), (64,64,8000), (96,112,4000), (224,224,3000), (16,64,5000), (128,256,2000)]:
    c = surface_voxels(rng, n); n = c.shape[0]
    nbr = oracle.kernel_map(c, c, 3, 1)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) / np.sqrt(27*cin)).astype(np.float32)
    g = rng.standard_normal((n, cout)).astype(np.float32)
    want = oracle.conv_fwd(x, w, nbr, n)
    want_gin, _ = oracle.conv_bwd(x, w, g, nbr)
    dn = D(nbr)
    res = []
    for algo in (2, 3):
        y = ops.conv_table(D(x), D(w), dn, n, 27, cin, cout, algo=algo)
        gi = ops.conv_table(D(g), D(w), dn, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=algo)
        torch.cuda.synchronize()
        res.append((rel(y, want), rel(gi, want_gin)))
    ys = ops.conv_table(D(x), D(w), dn, n, 27, cin, cout, algo=1)
    print(cin, cout, n, '3xtf32 fwd/dgrad %.2e %.2e | tf32 %.2e %.2e | simt %.2e' % (res[0] + res[1] + (rel(ys, want),)), flush=True)
print("== 1x1 / identity")
x = rng.standard_normal((5000, 32)).astype(np.float32); w = rng.standard_normal((32, 16)).astype(np.float32)
y = ops.conv_table(D(x), D(w), None, 5000, 1, 32, 16, algo=2); print('1x1', rel(y, x.astype(np.float64) @ w.astype(np.float64)))
print("== pairs mode (strided/transposed)")
for (cin, cout) in [(16, 32), (32, 48), (112, 96)]:
    c = surface_voxels(rng, 15000); n = c.shape[0]
    _, _, oc = oracle.coord_unique(c, 2); m = oc.shape[0]
    nbr_down = oracle.kernel_map(c, oc, 2, 1)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    wd = (rng.standard_normal((8, cin, cout)) / np.sqrt(8*cin)).astype(np.float32)
    wu = (rng.standard_normal((8, cout, cin)) / np.sqrt(8*cout)).astype(np.float32)
    want_y = oracle.conv_fwd(x, wd, nbr_down, m)
    want_z = oracle.convT_fwd(want_y, wu, nbr_down, n)
    dn = D(nbr_down)
    pin, pout, koff, _ = ops.pairs_from_nbr(dn)
    y = ops.conv_table(D(x), D(wd), dn, m, 8, cin, cout, algo=2)
    z = ops.conv_pairs(D(want_y), D(wu), pout, pin, koff, n, 8, cout, cin, n, algo=2)
    torch.cuda.synchronize()
    print(cin, cout, 'down %.2e up %.2e' % (rel(y, want_y), rel(z, want_z)), flush=True)
print("== timing level-0-like 16->16, 325k rows")
c = surface_voxels(rng, 330000, batch=4); n = c.shape[0]
table, _, _, oc = ops.coord_unique(D(c), 1); nbr = ops.kernel_map(oc, table, 3, 1)
for (cin, cout) in [(16,16),(32,32),(64,64)]:
    x = torch.randn(n, cin, device=dev); w = torch.randn(27, cin, cout, device=dev) * 0.05
    for algo in (1, 2, 3):
        for _ in range(3): ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=algo)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=algo)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
        print('rows', n, cin, cout, 'algo', algo, '%.1f us' % (dt * 1e6), flush=True)
