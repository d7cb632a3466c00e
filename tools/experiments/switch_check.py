"""Compact correctness check of the kernels behind the experiment switches (run under B2S_* environment variables by
tests/test_gpu_switches.py): table convolutions on a sorted large map / a small split map / a tiny wide map vs the fp32
FMA path, the weight gradient vs an fp64 reference, BatchNorm forward + backward vs torch.  Prints SWITCH_CHECK_OK."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops
rng = np.random.default_rng(0)
torch.manual_seed(0)
def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for rows, cin, cout in ((40_000, 16, 16), (40_000, 32, 16), (9_000, 64, 64), (300, 224, 112)):
    co = torch.from_numpy(surface_voxels(rng, rows, batch=4)).cuda()
    table, _, _, oc = ops.coord_unique(co, 1)
    n = oc.size(0)
    nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
    x = torch.randn(n, cin, device="cuda"); w = torch.randn(27, cin, cout, device="cuda") * 0.05
    g = torch.randn(n, cout, device="cuda"); sc = torch.randn(n, cout, device="cuda")
    ref = ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=1) + sc
    refg = ops.conv_table(g, w, nbr, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=1)
    variants = [(nbr, dict(tile_mask=tmask))]
    if n >= 32768:
        perm, nbs, tms = ops.tile_order(nbr)
        variants.append((nbs, dict(tile_mask=tms, out_rows=perm)))
    for tb, kw in variants:
        y = ops.conv_table(x, w, tb, n, 27, cin, cout, algo=0, add_src=sc, packed=ops.conv_pack(w), **kw)
        gi = ops.conv_table(g, w, tb, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=0, **kw)
        assert rel(y, ref) < 1e-4 and rel(gi, refg) < 1e-4, (rows, cin, cout, rel(y, ref), rel(gi, refg))
    pin, pout, koff, _ = ops.pairs_from_nbr(nbr)
    gw = ops.conv_wgrad(x, g, pin, pout, koff, 27, cin, cout, n * 27)
    for k in (0, 13, 26):
        sel = nbr[:, k] >= 0
        want = x[nbr[sel, k].long()].double().T @ g[sel].double()
        assert rel(gw[k].double(), want) < 1e-4, (rows, cin, cout, k)
for n, c in ((70_000, 16), (900, 64)):
    x = (torch.randn(n, c, device="cuda") * 2 + 1).requires_grad_()
    bn = torch.nn.BatchNorm1d(c).cuda()
    dy = torch.randn(n, c, device="cuda")
    torch.relu(bn(x)).backward(dy)
    y, mean, rstd = ops.bn_forward(x.detach(), bn.eps, 0.0, None, None, bn.weight.detach(), bn.bias.detach(), True)
    res = ops.bn_backward(x.detach(), y, dy, mean, rstd, bn.weight.detach(), True, True)
    assert rel(y, torch.relu(bn(x)).detach()) < 1e-5
    assert rel(res[0], x.grad) < 1e-4, rel(res[0], x.grad)
torch.cuda.synchronize()
print("SWITCH_CHECK_OK")
