"""Round-2 check of the fused BN->ReLU->(de)conv path (harness/models.py: FUSED_UPDOWN; csrc/fused.cu: b2s_bnconv_*).
Run on the GPU box: compares a TinyUnet forward/backward with the switch off and on (features, input gradient,
BatchNorm gradients and running statistics must be torch.equal; weight gradients within 1e-5) and prints the step time.
NOT verified on a GPU yet.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from minsu3d_b200 import MinkowskiEngine as ME
from minsu3d_b200.harness import models, scenes

batch = scenes.make_batch([11, 12], "cuda", n_points=30_000)
coords = batch["voxel_xyz"][:20000].contiguous()
torch.manual_seed(0)
net = models.TinyUnet(16).cuda().train()
state = {k: v.clone() for k, v in net.state_dict().items()}
x0 = torch.randn(coords.size(0), 16, device="cuda")
g = None
res = []
for fused in (False, True):
    models.FUSED_UPDOWN = fused
    net.load_state_dict(state)
    net.zero_grad(set_to_none=True)
    xa = x0.clone().requires_grad_(True)
    y = net(ME.SparseTensor(features=xa, coordinates=coords)).F
    g = torch.randn_like(y) if g is None else g
    y.backward(g)
    res.append((y.detach().clone(), xa.grad.clone(), {k: p.grad.clone() for k, p in net.named_parameters()},
                {k: v.clone() for k, v in net.state_dict().items()}))
(ya, gxa, ga, sa), (yb, gxb, gb, sb) = res
print("features equal", torch.equal(ya, yb), "input grad equal", torch.equal(gxa, gxb))
for k in ga:
    if k.endswith("kernel"):
        err = float((ga[k] - gb[k]).abs().max() / ga[k].abs().max().clamp_min(1e-12))
        assert err < 1e-5, (k, err)
    else:
        assert torch.equal(ga[k], gb[k]), k
for k in sa:
    if not k.endswith("kernel"):
        assert torch.equal(sa[k], sb[k]), k
assert torch.equal(ya, yb) and torch.equal(gxa, gxb)
print("fused up/down path OK")
