"""Round-2 check + timing of the persistent tcgen05 convolution (csrc/conv_tcp.cu) against the fp32 FMA path.
    B2S_TC_PERSIST=1 python tools/experiments/tcp_check.py     # persistent kernel (default)
    B2S_TC_PERSIST=0 python tools/experiments/tcp_check.py     # round-1 per-tile kernel, same calls
Shapes: the U-Net levels of the benchmark batch (rows x channels), 1x1 shortcuts, a strided 2^3 table, the residual add,
pre-packed weights.  Prints max relative error vs algo=1 and the time per launch (warm L2, packed weights cached)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
import torch
from helpers import surface_voxels
from minsu3d_b200 import ops

print("B2S_TC_PERSIST =", os.environ.get("B2S_TC_PERSIST", "1"), flush=True)
rng = np.random.default_rng(0)
bad = 0


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for rows, cin, cout in ((330_000, 16, 16), (330_000, 32, 16), (85_000, 32, 32), (85_000, 64, 32), (22_000, 48, 48),
                        (9_000, 64, 64), (2_100, 80, 80), (500, 96, 96), (100, 112, 112), (100, 224, 112),
                        (9_000, 224, 224), (40_000, 16, 16), (130, 16, 16), (128, 32, 32)):
    co = surface_voxels(rng, rows, batch=4)
    n = co.shape[0]
    table, _, _, oc = ops.coord_unique(torch.from_numpy(co).cuda(), 1)
    nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
    x = torch.randn(n, cin, device="cuda")
    w = torch.randn(27, cin, cout, device="cuda") * 0.05
    sc = torch.randn(n, cout, device="cuda")
    ref = ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=1)
    packed = ops.conv_pack(w)
    variants = [("row-order", dict(tile_mask=tmask))]
    if n >= 32768:
        perm, nbs, tms = ops.tile_order(nbr)
        variants.append(("sorted", dict(tile_mask=tms, out_rows=perm)))
    for name, kw in variants:
        tb = nbs if "out_rows" in kw else nbr
        for algo in (2, 3):
            y = ops.conv_table(x, w, tb, n, 27, cin, cout, algo=algo, packed=packed, **kw)
            y2 = ops.conv_table(x, w, tb, n, 27, cin, cout, algo=algo, **kw)  # packs per call
            ya = ops.conv_table(x, w, tb, n, 27, cin, cout, algo=algo, packed=packed, add_src=sc, **kw)
            err = float((y - ref).abs().max() / ref.abs().max())
            erra = float((ya - (ref + sc)).abs().max() / (ref + sc).abs().max())
            same = bool(torch.equal(y, y2))
            tol = 1e-4 if algo == 2 else 5e-3
            ok = err < tol and erra < tol and same
            bad += not ok
            t = timeit(lambda: ops.conv_table(x, w, tb, n, 27, cin, cout, algo=algo, packed=packed, **kw))
            print("%-9s rows %6d %3d->%3d algo %d  %7.1f us  err %.1e  +res %.1e  packed==unpacked %s  %s" % (
                name, n, cin, cout, algo, t, err, erra, same, "ok" if ok else "FAIL"), flush=True)
    # data gradient (reversed offsets, transposed weights)
    g = torch.randn(n, cout, device="cuda")
    refg = ops.conv_table(g, w, nbr, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=1)
    yg = ops.conv_table(g, w, nbr, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=2, tile_mask=tmask,
                        packed=packed)
    errg = float((yg - refg).abs().max() / refg.abs().max())
    bad += not errg < 1e-4
    print("dgrad     rows %6d %3d->%3d          err %.1e  %s" % (n, cout, cin, errg, "ok" if errg < 1e-4 else "FAIL"), flush=True)

# 1x1 convolution (identity map) and a strided 2^3 table
for rows, cin, cout in ((330_000, 32, 16), (9_000, 128, 64), (77, 224, 112)):
    x = torch.randn(rows, cin, device="cuda")
    w = torch.randn(cin, cout, device="cuda") * 0.1
    ref = x @ w
    y = ops.conv_table(x, w, None, rows, 1, cin, cout, algo=2, packed=ops.conv_pack(w))
    err = float((y - ref).abs().max() / ref.abs().max())
    t = timeit(lambda: ops.conv_table(x, w, None, rows, 1, cin, cout, algo=2, packed=ops.conv_pack(w)))
    bad += not err < 1e-4
    print("1x1       rows %6d %3d->%3d  %7.1f us (incl. pack)  err %.1e  %s" % (rows, cin, cout, t, err, "ok" if err < 1e-4 else "FAIL"),
          flush=True)
co = surface_voxels(rng, 85_000, batch=4)
table, _, _, oc = ops.coord_unique(torch.from_numpy(co).cuda(), 1)
t2, _, _, oc2 = ops.coord_unique(oc, 2)
nbr8, tm8 = ops.kernel_map(oc2, table, 2, 1, with_tile_mask=True)
x = torch.randn(oc.size(0), 32, device="cuda")
w = torch.randn(8, 32, 48, device="cuda") * 0.1
ref = ops.conv_table(x, w, nbr8, oc2.size(0), 8, 32, 48, algo=1)
y = ops.conv_table(x, w, nbr8, oc2.size(0), 8, 32, 48, algo=2, tile_mask=tm8, packed=ops.conv_pack(w))
err = float((y - ref).abs().max() / ref.abs().max())
bad += not err < 1e-4
print("strided   rows %6d  32-> 48  err %.1e  %s" % (oc2.size(0), err, "ok" if err < 1e-4 else "FAIL"), flush=True)
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
