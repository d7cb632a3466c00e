"""Forward AND backward of the MinkUNet backbone on the oracle's sparse ops (CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the model-level gradient checker of tests/ at the benchmark
configuration.  PARITY UNPINNED against the MinkowskiEngine binary (un-vendored dependency, see oracle/__init__.py);
the sparse products are oracle.c's restatement (`orc_conv_fwd`, `orc_conv_bwd`, `orc_convT_fwd`).

Structure follows minsu3d/model/module/backbone.py:8-43 and common.py:21-95 layer by layer (same walk as
oracle/me_unet.py).  The sparse convolutions are torch.autograd.Functions whose forward / backward call the C
oracle; BatchNorm (batch statistics, float64), ReLU, concatenation, the devoxelise gather and the two linear heads
are torch CPU ops, so torch's autograd chains everything and `.grad` of every parameter is the oracle gradient.
"""
import numpy as np
import torch

from . import conv_bwd, conv_fwd, convT_fwd
from .me_unet import _Maps


class _Conv(torch.autograd.Function):
    """out[o] = sum_k x[nbr[o,k]] @ W[k]  (appendix A.7); backward = orc_conv_bwd."""

    @staticmethod
    def forward(ctx, x, W, nbr, n_out):
        ctx.save_for_backward(x, W)
        ctx.nbr = nbr
        return torch.from_numpy(conv_fwd(x.detach().numpy(), W.detach().numpy(), nbr, n_out))

    @staticmethod
    def backward(ctx, gout):
        x, W = ctx.saved_tensors
        gin, gW = conv_bwd(x.detach().numpy(), W.detach().numpy(), gout.contiguous().numpy(), ctx.nbr)
        return torch.from_numpy(gin), torch.from_numpy(gW), None, None


class _ConvT(torch.autograd.Function):
    """Transposed 2^3/stride-2 convolution on the forward strided map nbr_down[coarse, k] = fine row:
    out[f] = x[parent(f)] @ W[k(f)].  Backward through the same restated ops:
    gx[c] = sum_k gout[nbr_down[c,k]] @ W[k]^T (a table convolution with transposed weights) and
    gW[k] = x[c]^T @ gout[nbr_down[c,k]] (orc_conv_bwd with the operands' roles swapped, then transposed)."""

    @staticmethod
    def forward(ctx, x, W, nbr_down, n_fine):
        ctx.save_for_backward(x, W)
        ctx.nbr = nbr_down
        return torch.from_numpy(convT_fwd(x.detach().numpy(), W.detach().numpy(), nbr_down, n_fine))

    @staticmethod
    def backward(ctx, gout):
        x, W = ctx.saved_tensors
        Wt = np.ascontiguousarray(W.detach().numpy().transpose(0, 2, 1))
        g = gout.contiguous().numpy()
        gx = conv_fwd(g, Wt, ctx.nbr, x.shape[0])
        _, gWt = conv_bwd(g, Wt, x.detach().numpy(), ctx.nbr)
        return torch.from_numpy(gx), torch.from_numpy(np.ascontiguousarray(gWt.transpose(0, 2, 1))), None, None


def _bn_relu(x, sd, prefix, relu=True, eps=1e-5):
    y = torch.nn.functional.batch_norm(x.double(), None, None, sd[prefix + ".weight"].double(),
                                       sd[prefix + ".bias"].double(), True, 0.0, eps).float()
    return torch.relu(y) if relu else y


def _residual(x, sd, p, maps, ts):
    shortcut = x
    if (p + ".downsample.0.kernel") in sd:
        shortcut = x @ sd[p + ".downsample.0.kernel"]
    nbr = maps.same(ts)
    y = _Conv.apply(_bn_relu(x, sd, p + ".conv_branch.0.bn"), sd[p + ".conv_branch.2.kernel"], nbr, x.shape[0])
    y = _Conv.apply(_bn_relu(y, sd, p + ".conv_branch.3.bn"), sd[p + ".conv_branch.5.kernel"], nbr, x.shape[0])
    return y + shortcut


def _ublock(x, sd, p, maps, ts, depth, reps):
    for i in range(reps):
        x = _residual(x, sd, "%s.blocks.block%d" % (p, i), maps, ts)
    if depth > 1:
        skip = x
        nbr_down = maps.down(ts)
        n_coarse = maps.coords[2 * ts].shape[0]
        y = _Conv.apply(_bn_relu(x, sd, p + ".conv.0.bn"), sd[p + ".conv.2.kernel"], nbr_down, n_coarse)
        y = _ublock(y, sd, p + ".u", maps, 2 * ts, depth - 1, reps)
        y = _ConvT.apply(_bn_relu(y, sd, p + ".deconv.0.bn"), sd[p + ".deconv.2.kernel"], nbr_down, x.shape[0])
        x = torch.cat((skip, y), dim=1)
        for i in range(reps):
            x = _residual(x, sd, "%s.blocks_tail.block%d" % (p, i), maps, ts)
    return x


def _head(x, sd, p):
    y = torch.nn.functional.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
    y = _bn_relu(y, sd, p + ".1")
    return torch.nn.functional.linear(y, sd[p + ".3.weight"], sd[p + ".3.bias"])


def backbone_forward(params, voxel_features, voxel_coords, v2p_map, depth=7, reps=2):
    """params: name -> float32 CPU tensor (leaves with requires_grad=True for the ones to differentiate), names as in
    the reference's state dict.  Returns torch tensors (point_features, semantic_scores, point_offsets)."""
    sd = params
    maps = _Maps(np.ascontiguousarray(voxel_coords, np.int32))
    x = torch.as_tensor(np.ascontiguousarray(voxel_features, np.float32))
    x = _Conv.apply(x, sd["backbone.unet.0.kernel"], maps.same(1), x.shape[0])
    x = _ublock(x, sd, "backbone.unet.1", maps, 1, depth, reps)
    x = _bn_relu(x, sd, "backbone.unet.2.bn")
    pf = x[torch.as_tensor(np.asarray(v2p_map)).long()]
    return {"point_features": pf, "semantic_scores": _head(pf, sd, "backbone.semantic_branch"),
            "point_offsets": _head(pf, sd, "backbone.offset_branch")}


def backbone_gradients(module, voxel_features, voxel_coords, v2p_map, g_sem, g_off, depth=7, reps=2):
    """Oracle gradients of  L = <semantic_scores, g_sem> + <point_offsets, g_off>  w.r.t. every backbone parameter
    of `module` (a harness / reference model whose state dict uses the reference's names).
    Returns (outputs dict of numpy arrays, grads dict name -> numpy array)."""
    params = {}
    for k, v in module.state_dict().items():
        if k.startswith("backbone.") and v.is_floating_point():
            params[k] = v.detach().cpu().float().clone()
    for k, p in module.named_parameters():
        if k in params:
            params[k].requires_grad_(True)
    out = backbone_forward(params, voxel_features, voxel_coords, v2p_map, depth, reps)
    loss = (out["semantic_scores"] * torch.as_tensor(g_sem)).sum() + (out["point_offsets"] * torch.as_tensor(g_off)).sum()
    loss.backward()
    grads = {k: v.grad.numpy() for k, v in params.items() if v.requires_grad and v.grad is not None}
    return {k: v.detach().numpy() for k, v in out.items()}, grads
