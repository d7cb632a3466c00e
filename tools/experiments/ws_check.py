"""Warp-stream narrow-layer convolution (csrc/conv_ws.cu) vs the fp32 FMA path: error and time per launch (warm L2).
    B2S_CONV_WS=1 (default) / 0 (persistent tcgen05 kernel, same calls)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np, torch
from helpers import surface_voxels
from minsu3d_b200 import ops
ALGO = 4 if os.environ.get("B2S_CONV_WS", "1") != "0" else 2
print("algo", ALGO, "(4 = warp-stream, 2 = persistent tcgen05)", flush=True)
rng = np.random.default_rng(0)
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
from minsu3d_b200.harness import scenes
batch = scenes.make_batch([0, 1, 2, 3], "cuda", 100_000)
cases = [("bench level 0", batch["voxel_xyz"])]
for rows in (90_000, 30_000, 3_000, 130):
    cases.append(("%d rows" % rows, torch.from_numpy(surface_voxels(rng, rows, batch=4)).cuda()))
for name, co in cases:
    table, _, _, oc = ops.coord_unique(co, 1)
    n = oc.size(0)
    nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
    variants = [("row-order", nbr, dict(tile_mask=tmask))]
    if n >= 32768:
        perm, nbs, tms = ops.tile_order(nbr)
        variants.append(("sorted", nbs, dict(tile_mask=tms, out_rows=perm)))
    for cin, cout in ((16, 16), (32, 16), (16, 32), (32, 32)):
        x = torch.randn(n, cin, device="cuda"); w = torch.randn(27, cin, cout, device="cuda") * 0.05
        g = torch.randn(n, cout, device="cuda"); sc = torch.randn(n, cout, device="cuda")
        ref = ops.conv_table(x, w, nbr, n, 27, cin, cout, algo=1)
        refg = ops.conv_table(g, w, nbr, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=1)
        packed = ops.conv_pack(w)
        for vn, tb, kw in variants:
            y = ops.conv_table(x, w, tb, n, 27, cin, cout, algo=ALGO, packed=packed, add_src=sc, **kw)
            gi = ops.conv_table(g, w, tb, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=ALGO, packed=packed, **kw)
            e1 = float((y - (ref + sc)).abs().max() / (ref + sc).abs().max()); e2 = float((gi - refg).abs().max() / refg.abs().max())
            t1 = timeit(lambda: ops.conv_table(x, w, tb, n, 27, cin, cout, algo=ALGO, packed=packed, **kw))
            t2 = timeit(lambda: ops.conv_table(g, w, tb, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=ALGO, packed=packed, **kw))
            print("%-14s %7d rows %2d->%2d %-9s fwd %6.1f us (err %.1e)  dgrad %6.1f us (err %.1e) %s" % (
                name, n, cin, cout, vn, t1, e1, t2, e2, "" if max(e1, e2) < 1e-4 else "  <-- FAIL"))
