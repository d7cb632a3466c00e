"""ncu target: weight-gradient kernels on the level-0 map (16x16) and a level-3-like map (64x64)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
import torch
from helpers import surface_voxels
from minsu3d_b200 import ops
from minsu3d_b200.harness import scenes

batch = scenes.make_batch([0, 1, 2, 3], "cuda", 100_000)
table, _, _, oc = ops.coord_unique(batch["voxel_xyz"], 1)
cases = [(oc, table, 16)]
rng = np.random.default_rng(0)
co = torch.from_numpy(surface_voxels(rng, 10_000, batch=4)).cuda()
t2, _, _, oc2 = ops.coord_unique(co, 1)
cases.append((oc2, t2, 64))
work = []
for c_, t_, ch in cases:
    nbr = ops.kernel_map(c_, t_, 3, 1)
    pin, pout, koff, _ = ops.pairs_from_nbr(nbr)
    n = c_.size(0)
    x = torch.randn(n, ch, device="cuda")
    g = torch.randn(n, ch, device="cuda")
    work.append((x, g, pin, pout, koff, ch, n))
for w in work:
    for _ in range(2):
        ops.conv_wgrad(w[0], w[1], w[2], w[3], w[4], 27, w[5], w[5], w[6] * 27)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for w in work:
    ops.conv_wgrad(w[0], w[1], w[2], w[3], w[4], 27, w[5], w[5], w[6] * 27)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
