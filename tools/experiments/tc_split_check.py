"""Round-2 experiment: offset split of the tcgen05 conv for maps with few tiles (conv_tc.cu, B2S_TC_SPLIT=1).

Run on the GPU box, once per mode (the switch is read once per process):
    B2S_TC_SPLIT=0 python tools/experiments/tc_split_check.py
    B2S_TC_SPLIT=1 python tools/experiments/tc_split_check.py
Prints, for the deep U-Net levels of the benchmark batch (rows x channels), the max relative error against the fp32 FMA
path and the time per launch.  Expected from the launch list (profiles/r01_step_launches_final2.txt): levels 2-6 cost
30-60 us per conv unsplit (80-190 slabs walked serially by 2-170 CTAs); the split should bring them to 10-20 us.
NOT verified on a GPU yet (written after the round-1 GPU budget was spent).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
import torch
from helpers import surface_voxels
from minsu3d_b200 import ops

print("B2S_TC_SPLIT =", os.environ.get("B2S_TC_SPLIT", "0"))
rng = np.random.default_rng(0)
for rows, c in ((85_000, 32), (22_000, 48), (9_000, 64), (2_100, 80), (500, 96), (100, 112), (9_000, 224)):
    co = surface_voxels(rng, rows, batch=4)
    n = co.shape[0]
    table, _, _, oc = ops.coord_unique(torch.from_numpy(co).cuda(), 1)
    nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
    x = torch.randn(n, c, device="cuda")
    w = torch.randn(27, c, c, device="cuda") * 0.05
    ref = ops.conv_table(x, w, nbr, n, 27, c, c, algo=1)
    for _ in range(3):
        y = ops.conv_table(x, w, nbr, n, 27, c, c, algo=2, tile_mask=tmask)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.conv_table(x, w, nbr, n, 27, c, c, algo=2, tile_mask=tmask)
    e1.record()
    torch.cuda.synchronize()
    err = float((y - ref).abs().max() / ref.abs().max())
    y2 = ops.conv_table(x, w, nbr, n, 27, c, c, algo=2, tile_mask=tmask)
    print("rows %6d  c %3d  tiles %4d  %.1f us  err %.2e  deterministic %s" % (
        n, c, (n + 127) // 128, e0.elapsed_time(e1) / 20 * 1e3, err, bool(torch.equal(y, y2))), flush=True)
