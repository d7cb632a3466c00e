"""ncu target: the level-0 convolution of the benchmark batch (16 -> 16, 3^3, mask-sorted tiles, packed weights cached)
on the persistent tcgen05 kernel; 3 warm-up launches, then profiler range around 2 launches."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from minsu3d_b200 import ops
from minsu3d_b200.harness import scenes

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 16
batch = scenes.make_batch([0, 1, 2, 3], "cuda", 100_000)
table, _, _, oc = ops.coord_unique(batch["voxel_xyz"], 1)
nbr, tmask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
perm, nbs, tms = ops.tile_order(nbr)
m = oc.size(0)
x = torch.randn(m, cin, device="cuda")
w = torch.randn(27, cin, cout, device="cuda") * 0.05
packed = ops.conv_pack(w)
for _ in range(3):
    ops.conv_table(x, w, nbs, m, 27, cin, cout, algo=2, tile_mask=tms, out_rows=perm, packed=packed)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(2):
    ops.conv_table(x, w, nbs, m, 27, cin, cout, algo=2, tile_mask=tms, out_rows=perm, packed=packed)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
