"""The oracle's model-level backward (oracle/me_unet_grad.py: orc_conv_bwd + the transposed-convolution identities)
against an independent float64 torch-autograd restatement that uses nothing but index_add over explicit pair lists.
This is what makes the oracle usable as the gradient checker of the GPU path at the benchmark configuration."""
import numpy as np
import torch

import oracle
from oracle import me_unet, me_unet_grad


def _dense_autograd_grads(model, vf, vx, vm, gs, go, depth, reps=2):
    P = {k: v.detach().double().clone() for k, v in model.state_dict().items() if k.startswith("backbone.")
         and v.is_floating_point()}
    names = {k for k, _ in model.named_parameters()}
    for k in P:
        if k in names:
            P[k].requires_grad_(True)
    maps = me_unet._Maps(vx)

    def conv(x, w, nbr, n_out, transpose=False):
        out = torch.zeros(n_out, w.shape[-1], dtype=torch.float64)
        for k in range(nbr.shape[1]):
            o = np.nonzero(nbr[:, k] >= 0)[0]
            if o.size == 0:
                continue
            i = torch.from_numpy(nbr[o, k].astype(np.int64))
            o = torch.from_numpy(o.astype(np.int64))
            out = out.index_add(0, i, x[o] @ w[k]) if transpose else out.index_add(0, o, x[i] @ w[k])
        return out

    def bnr(x, p):
        return torch.relu(torch.nn.functional.batch_norm(x, None, None, P[p + ".weight"], P[p + ".bias"], True, 0.0, 1e-5))

    def res(x, p, ts):
        sc = x @ P[p + ".downsample.0.kernel"] if (p + ".downsample.0.kernel") in P else x
        nbr = maps.same(ts)
        y = conv(bnr(x, p + ".conv_branch.0.bn"), P[p + ".conv_branch.2.kernel"], nbr, x.shape[0])
        y = conv(bnr(y, p + ".conv_branch.3.bn"), P[p + ".conv_branch.5.kernel"], nbr, x.shape[0])
        return y + sc

    def ublock(x, p, ts, d):
        for i in range(reps):
            x = res(x, "%s.blocks.block%d" % (p, i), ts)
        if d > 1:
            skip = x
            nd = maps.down(ts)
            y = conv(bnr(x, p + ".conv.0.bn"), P[p + ".conv.2.kernel"], nd, maps.coords[2 * ts].shape[0])
            y = ublock(y, p + ".u", 2 * ts, d - 1)
            y = conv(bnr(y, p + ".deconv.0.bn"), P[p + ".deconv.2.kernel"], nd, x.shape[0], transpose=True)
            x = torch.cat((skip, y), 1)
            for i in range(reps):
                x = res(x, "%s.blocks_tail.block%d" % (p, i), ts)
        return x

    def head(x, p):
        y = torch.nn.functional.linear(x, P[p + ".0.weight"], P[p + ".0.bias"])
        y = bnr(y, p + ".1")
        return torch.nn.functional.linear(y, P[p + ".3.weight"], P[p + ".3.bias"])

    x = conv(torch.from_numpy(vf).double(), P["backbone.unet.0.kernel"], maps.same(1), vf.shape[0])
    x = bnr(ublock(x, "backbone.unet.1", 1, depth), "backbone.unet.2.bn")
    pf = x[torch.from_numpy(vm).long()]
    loss = (head(pf, "backbone.semantic_branch") * torch.from_numpy(gs).double()).sum() + \
           (head(pf, "backbone.offset_branch") * torch.from_numpy(go).double()).sum()
    loss.backward()
    return {k: v.grad.numpy() for k, v in P.items() if v.requires_grad}


def test_oracle_unet_gradients_match_dense_float64_autograd():
    from minsu3d_b200.harness import models, scenes
    torch.manual_seed(1)
    depth = 3
    model = models.build_model(models.Config.for_model("pointgroup", blocks=[1, 2, 3]))
    b = scenes.collate([scenes.make_scene(3, 4000)], "cpu")
    vf, vx, vm = b["voxel_features"].numpy(), b["voxel_xyz"].numpy(), b["voxel_point_map"].numpy()
    rng = np.random.default_rng(0)
    gs = rng.normal(size=(vm.shape[0], 20)).astype(np.float32)
    go = rng.normal(size=(vm.shape[0], 3)).astype(np.float32)
    out, grads = me_unet_grad.backbone_gradients(model, vf, vx, vm, gs, go, depth=depth)
    want = _dense_autograd_grads(model, vf, vx, vm, gs, go, depth)
    assert set(grads) == set(want) and len(grads) > 40
    for k, g in grads.items():
        if k.endswith("_branch.0.bias"):  # a bias in front of a BatchNorm: the true gradient is exactly zero
            assert np.linalg.norm(g) < 1e-3 and np.linalg.norm(want[k]) < 1e-9, k
            continue
        err = np.linalg.norm(g - want[k]) / np.linalg.norm(want[k])
        assert err < 1e-5, "%s: rel l2 %.3e" % (k, err)
    # and the forward of the autograd oracle is the numpy oracle (same C calls)
    ref = me_unet.backbone_forward(me_unet.numpy_state_dict(model), vf, vx, vm, depth=depth)
    for k in ref:
        assert np.array_equal(out[k], ref[k]), k
