"""ncu target: weight-gradient kernel on chosen shapes.  usage: wgrad_ncu_target.py rows:ca:cg [rows:ca:cg ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
import torch
from helpers import surface_voxels
from minsu3d_b200 import ops

rng = np.random.default_rng(0)
work = []
for spec in sys.argv[1:] or ["330000:16:16", "25000:64:64"]:
    rows, ca, cg = (int(v) for v in spec.split(":"))
    co = torch.from_numpy(surface_voxels(rng, rows, batch=4)).cuda()
    t2, _, _, oc2 = ops.coord_unique(co, 1)
    nbr = ops.kernel_map(oc2, t2, 3, 1)
    pin, pout, koff, _ = ops.pairs_from_nbr(nbr)
    n = oc2.size(0)
    work.append((torch.randn(n, ca, device="cuda"), torch.randn(n, cg, device="cuda"), pin, pout, koff, ca, cg, n))
for w in work:
    for _ in range(2):
        ops.conv_wgrad(w[0], w[1], w[2], w[3], w[4], 27, w[5], w[6], w[7] * 27)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for w in work:
    ops.conv_wgrad(w[0], w[1], w[2], w[3], w[4], 27, w[5], w[6], w[7] * 27)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
