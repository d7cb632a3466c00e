"""ncu target: one-launch BatchNorm forward / backward on the level-0 tensor of the benchmark batch ([325422, 16])."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from minsu3d_b200 import ops
n, c = 325422, 16
x = torch.randn(n, c, device="cuda"); dy = torch.randn(n, c, device="cuda")
g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
for _ in range(3):
    y, mean, rstd = ops.bn_forward(x, 1e-4, 0.1, rm, rv, g, b, True)
    ops.bn_backward(x, y, dy, mean, rstd, g, True, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
y, mean, rstd = ops.bn_forward(x, 1e-4, 0.1, rm, rv, g, b, True)
ops.bn_backward(x, y, dy, mean, rstd, g, True, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
