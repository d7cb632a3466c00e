"""GPU model-level parity and end-to-end checks (BASELINE.json configs 1-4 at test sizes)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import me_unet

pytestmark = pytest.mark.gpu


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)


def _rel_l2(got, want):
    """Relative L2 error: robust to the handful of ReLU pre-activations within fp32 rounding of zero, whose
    mask may legitimately differ between the fp32 kernels and the float64 reference."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-12)


@pytest.fixture(scope="module")
def batch():
    from minsu3d_b200.harness import scenes
    return scenes.make_batch([11, 12], "cuda", n_points=30_000)


def test_backbone_forward_matches_cpu_restatement(batch):
    """config 1 at test size: MinkUNet (m=16, 7 levels) forward, CUDA vs the restated ME CPU backend."""
    from minsu3d_b200.harness import models
    torch.manual_seed(123)
    model = models.build_model(models.Config.for_model("pointgroup")).cuda().train()
    out = model.backbone_forward(batch)
    sd = me_unet.numpy_state_dict(model)
    want = me_unet.backbone_forward(sd, batch["voxel_features"].cpu().numpy(), batch["voxel_xyz"].cpu().numpy(),
                                    batch["voxel_point_map"].cpu().numpy())
    for k in ("point_features", "semantic_scores", "point_offsets"):
        err = _rel(out[k].detach().cpu().numpy(), want[k])
        assert err < 1e-4, "%s rel err %.3e" % (k, err)


def test_backbone_gradients_match_dense_torch_autograd(batch):
    """Gradients of a 2-level TinyUnet through libb2s vs the same net built from torch index ops."""
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200.harness import models
    torch.manual_seed(5)
    net = models.TinyUnet(16).cuda().train()
    coords = batch["voxel_xyz"][:12000].contiguous()
    x0 = torch.randn(coords.size(0), 16, device="cuda")
    xa = x0.clone().requires_grad_(True)
    ya = net(ME.SparseTensor(features=xa, coordinates=coords)).F
    g = torch.randn_like(ya)
    ya.backward(g)
    got = {n: p.grad.detach().cpu().numpy() for n, p in net.named_parameters()}
    got_x = xa.grad.cpu().numpy()

    # reference: float64 torch autograd over explicit pair lists from the oracle's kernel maps
    c = coords.cpu().numpy()
    nbr3 = {1: oracle.kernel_map(c, c, 3, 1)}
    _, _, c2 = oracle.coord_unique(c, 2)
    nbr3[2] = oracle.kernel_map(c2, c2, 3, 2)
    nbr_down = oracle.kernel_map(c, c2, 2, 1)
    params = {n: p.detach().double().clone().requires_grad_(True) for n, p in net.named_parameters()}

    def conv(x, w, nbr, n_out, transpose=False):
        out = torch.zeros(n_out, w.shape[-1], dtype=torch.float64, device=x.device)
        for k in range(nbr.shape[1]):
            o = np.nonzero(nbr[:, k] >= 0)[0]
            if o.size == 0:
                continue
            i = torch.from_numpy(nbr[o, k]).long().cuda()
            o = torch.from_numpy(o).long().cuda()
            if transpose:
                out = out.index_add(0, i, x[o] @ w[k])
            else:
                out = out.index_add(0, o, x[i] @ w[k])
        return out

    def bn_relu(x, p):
        y = torch.nn.functional.batch_norm(x, None, None, params[p + ".bn.weight"], params[p + ".bn.bias"], True, 0.1, 1e-5)
        return torch.relu(y)

    def res(x, p, ts):
        n = x.size(0)
        sc = x @ params[p + ".downsample.0.kernel"] if (p + ".downsample.0.kernel") in params else x
        y = conv(bn_relu(x, p + ".conv_branch.0"), params[p + ".conv_branch.2.kernel"], nbr3[ts], n)
        y = conv(bn_relu(y, p + ".conv_branch.3"), params[p + ".conv_branch.5.kernel"], nbr3[ts], n)
        return y + sc

    xb = x0.double().clone().requires_grad_(True)
    h = res(res(xb, "unet.0.blocks.block0", 1), "unet.0.blocks.block1", 1)
    d = conv(bn_relu(h, "unet.0.conv.0"), params["unet.0.conv.2.kernel"], nbr_down, c2.shape[0])
    d = res(res(d, "unet.0.u.blocks.block0", 2), "unet.0.u.blocks.block1", 2)
    u = conv(bn_relu(d, "unet.0.deconv.0"), params["unet.0.deconv.2.kernel"], nbr_down, c.shape[0], transpose=True)
    t = torch.cat((h, u), 1)
    t = res(res(t, "unet.0.blocks_tail.block0", 1), "unet.0.blocks_tail.block1", 1)
    yb = bn_relu(t, "unet.1")
    assert _rel(ya.detach().cpu().numpy(), yb.detach().cpu().numpy()) < 1e-4
    yb.backward(g.double())
    assert _rel_l2(got_x, xb.grad.cpu().numpy()) < 1e-3
    for n, p in params.items():
        err = _rel_l2(got[n], p.grad.cpu().numpy())
        assert err < 1e-3, "%s grad rel l2 err %.3e" % (n, err)


@pytest.mark.parametrize("name", ["pointgroup", "hais", "softgroup"])
def test_train_step_runs_and_learns(name, batch):
    """configs 2-4 at test size: full step incl. clustering + refinement net; loss decreases on a fixed batch."""
    from minsu3d_b200.harness import models, train
    cfg = models.Config.for_model(name, proposal_source="gt_noise")
    tr = train.Trainer(cfg, "cuda")
    losses = [float(tr.step(batch)) for _ in range(6)]
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0]
    keys = set(tr.last_losses)
    assert {"semantic_loss", "offset_norm_loss", "offset_dir_loss"} <= keys
    assert keys & {"score_loss", "classification_loss"}, "clustering stage produced no proposals: %s" % keys


def test_pointgroup_proposals_match_oracle_pipeline(batch):
    """The clustering stage end to end (ball query -> BFS -> concat) against the CPU oracle."""
    from minsu3d_b200.harness import models
    torch.manual_seed(0)
    cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
    model = models.build_model(cfg).cuda().train()
    with torch.no_grad():
        out = model(batch)
    scores, pidx, poff = out["proposal_scores"]
    sem_scores, offsets = model._cluster_inputs(batch, out)
    preds = sem_scores.argmax(1).to(torch.int16)
    obj = model._object_points(preds).cpu().numpy()
    xyz = batch["point_xyz"].cpu().numpy()[obj]
    shifted = (batch["point_xyz"] + offsets)[torch.from_numpy(obj).cuda()].cpu().numpy()
    bidx = batch["vert_batch_ids"].cpu().numpy()[obj]
    offs = np.concatenate(([0], np.cumsum(np.bincount(bidx)))).astype(np.int32)
    lab = preds.cpu().numpy()[obj]
    sets = []
    for pts in (xyz, shifted):
        i, sl = oracle.ballquery(pts, bidx, offs, cfg.cluster_radius)
        ci, co = oracle.pg_bfs_cluster(lab, i, sl, cfg.cluster_npoint_thre)
        sets.append((obj[ci[:, 1]], co))
    want_pts = np.concatenate((sets[0][0], sets[1][0]))
    want_off = np.concatenate((sets[0][1], sets[1][1][1:] + sets[0][1][-1]))
    assert np.array_equal(poff.cpu().numpy(), want_off)
    assert np.array_equal(pidx[:, 1].cpu().numpy(), want_pts)
    assert want_off.size > 3


def test_reference_wrappers_run_unmodified_on_dropin(batch):
    """install_as_reference_modules(): `import COMMON_OPS` / `import MinkowskiEngine` resolve to this package and
    reference-style calls (allocate outputs, call COMMON_OPS.*) work."""
    import minsu3d_b200
    minsu3d_b200.install_as_reference_modules()
    import COMMON_OPS
    import MinkowskiEngine as ME
    x = torch.randn(100, 3, device="cuda")
    offs = torch.tensor([0, 40, 100], dtype=torch.int32, device="cuda")
    out = torch.zeros((2, 3), device="cuda")
    COMMON_OPS.sec_mean(x, offs, out, 2, 3)
    assert np.array_equal(out.cpu().numpy(), oracle.sec("mean", x.cpu().numpy(), offs.cpu().numpy()))
    st = ME.SparseTensor(features=batch["voxel_features"], coordinates=batch["voxel_xyz"])
    assert st.F.shape[0] == batch["voxel_xyz"].shape[0] and st.tensor_stride == [1, 1, 1]
    q = ME.utils.sparse_quantize(batch["voxel_xyz"].long(), batch["voxel_features"], return_index=True,
                                 return_inverse=True, device="cuda")
    assert q[2].dtype == torch.int64 and torch.equal(q[3], torch.arange(st.F.shape[0], device="cuda"))


@pytest.mark.parametrize("cin,cout,n_rows", [(16, 16, 12_000), (32, 16, 12_000), (48, 48, 3_000), (16, 16, 40_000)])
def test_fused_residual_block_equals_module_by_module(batch, monkeypatch, cin, cout, n_rows):
    """b2s_resblock_forward/backward (one call per block) vs the MinkowskiEngine-shaped module sequence:
    identical features, input gradients, BatchNorm gradients and running statistics; weight gradients to 1e-5
    (their atomic flush order varies run to run).  40k rows also crosses KernelMap.SORTED_MIN_ROWS."""
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200.harness import models
    coords = batch["voxel_xyz"][:n_rows].contiguous()
    n = coords.size(0)
    torch.manual_seed(cin * 100 + cout)
    x0 = torch.randn(n, cin, device="cuda")
    g = torch.randn(n, cout, device="cuda")
    blk = models.ResidualBlock(cin, cout, 3).cuda().train()
    state = {k: v.clone() for k, v in blk.state_dict().items()}
    res = []
    for fused in (False, True):
        monkeypatch.setattr(models, "FUSED_BLOCKS", fused)
        blk.load_state_dict(state)
        blk.zero_grad(set_to_none=True)
        xa = x0.clone().requires_grad_(True)
        st = ME.SparseTensor(features=xa, coordinates=coords)
        y = blk(st)
        assert y.coordinate_map_key == st.coordinate_map_key
        y.F.backward(g)
        res.append((y.F.detach().clone(), xa.grad.clone(), {k: p.grad.clone() for k, p in blk.named_parameters()},
                    {k: v.clone() for k, v in blk.state_dict().items()}))
    (ya, gxa, ga, sa), (yb, gxb, gb, sb) = res
    assert torch.equal(ya, yb)
    assert torch.equal(gxa, gxb)
    for k in ga:
        if k.endswith("kernel"):
            assert _rel(gb[k].cpu().numpy(), ga[k].cpu().numpy()) < 1e-5, k
        else:
            assert torch.equal(ga[k], gb[k]), k
    for k in sa:
        if not k.endswith("kernel"):
            assert torch.equal(sa[k], sb[k]), k  # running_mean / running_var / num_batches_tracked


def test_fused_blocks_fall_back_in_eval_mode(batch, monkeypatch):
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200.harness import models
    coords = batch["voxel_xyz"][:5000].contiguous()
    x = torch.randn(coords.size(0), 16, device="cuda")
    blk = models.ResidualBlock(16, 16, 3).cuda().eval()
    outs = []
    for fused in (False, True):
        monkeypatch.setattr(models, "FUSED_BLOCKS", fused)
        with torch.no_grad():
            outs.append(blk(ME.SparseTensor(features=x, coordinates=coords)).F)
    assert torch.equal(outs[0], outs[1])


def test_loader_size_hints_remove_host_reads_and_are_validated(batch):
    """SparseTensor(coordinates_unique=True, level_sizes=...) skips the host reads of the device-side counts; the
    claims are checked on the device: identical backbone output when they hold, ValueError when they do not."""
    from minsu3d_b200 import MinkowskiEngine as ME, ops
    from minsu3d_b200.harness import models, scenes
    torch.manual_seed(123)
    model = models.build_model(models.Config.for_model("pointgroup")).cuda().train()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    sizes = batch["voxel_level_sizes"]
    assert sorted(sizes) == [2, 4, 8, 16, 32, 64] and all(v > 0 for v in sizes.values())
    outs = []
    for hints in (None, sizes):
        model.load_state_dict(state)
        outs.append(model.backbone(batch["voxel_features"], batch["voxel_xyz"], batch["voxel_point_map"], hints))
        ops.run_deferred_checks()
    for k in ("point_features", "semantic_scores", "point_offsets"):
        assert torch.equal(outs[0][k], outs[1][k]), k
    # a wrong level size is detected on the device
    bad = dict(sizes)
    bad[4] -= 1
    model.backbone(batch["voxel_features"], batch["voxel_xyz"], batch["voxel_point_map"], bad)
    with pytest.raises(ValueError, match="claimed"):
        ops.run_deferred_checks()
    # duplicated coordinates passed as unique are detected too
    c = torch.cat((batch["voxel_xyz"][:100], batch["voxel_xyz"][:5])).contiguous()
    ME.SparseTensor(features=torch.zeros(105, 4, device="cuda"), coordinates=c, coordinates_unique=True)
    with pytest.raises(ValueError, match="claimed"):
        ops.run_deferred_checks()
    ops.run_deferred_checks()  # the list is empty again


def test_fused_bn_relu_updown_conv_equals_module_by_module(batch, monkeypatch):
    """b2s_bnconv_forward/backward (BN -> ReLU -> strided / transposed convolution of a U-Net level as one call,
    common.py:67-77) against the module-by-module path on a two-level U-Net: identical features, input gradient,
    BatchNorm gradients and running statistics (same kernels, same order); weight gradients to 1e-5."""
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200.harness import models
    coords = batch["voxel_xyz"][:20000].contiguous()
    torch.manual_seed(0)
    net = models.TinyUnet(16).cuda().train()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    x0 = torch.randn(coords.size(0), 16, device="cuda")
    g = None
    res = []
    for fused in (False, True):
        monkeypatch.setattr(models, "FUSED_UPDOWN", fused)
        net.load_state_dict(state)
        net.zero_grad(set_to_none=True)
        xa = x0.clone().requires_grad_(True)
        y = net(ME.SparseTensor(features=xa, coordinates=coords)).F
        g = torch.randn_like(y) if g is None else g
        y.backward(g)
        res.append((y.detach().clone(), xa.grad.clone(), {k: p.grad.clone() for k, p in net.named_parameters()},
                    {k: v.clone() for k, v in net.state_dict().items()}))
    (ya, gxa, ga, sa), (yb, gxb, gb, sb) = res
    assert torch.equal(ya, yb) and torch.equal(gxa, gxb)
    for k in ga:
        if k.endswith("kernel"):
            assert _rel(gb[k].cpu().numpy(), ga[k].cpu().numpy()) < 1e-5, k
        else:
            assert torch.equal(ga[k], gb[k]), k
    for k in sa:
        if not k.endswith("kernel"):
            assert torch.equal(sa[k], sb[k]), k


def test_packed_weight_cache_follows_fused_optimizer_steps(batch):
    """The tensor-core operand images of a convolution kernel are cached per parameter (ops.PackedWeights).  The fused
    Adam updates parameters without moving their version counters, so the cache must be invalidated by the optimizer
    step itself: after every step the module must convolve with the NEW weights."""
    from minsu3d_b200 import MinkowskiEngine as ME, ops
    coords = batch["voxel_xyz"][:8000].contiguous()
    x = torch.randn(coords.size(0), 16, device="cuda")
    conv = ME.MinkowskiConvolution(16, 32, kernel_size=3, dimension=3).cuda()
    for opt in (torch.optim.Adam(conv.parameters(), lr=0.05, fused=True), torch.optim.SGD(conv.parameters(), lr=0.5)):
        for _ in range(2):
            st = ME.SparseTensor(features=x, coordinates=coords)
            y = conv(st).F
            kmap = st.coordinate_manager.kernel_map(st.coordinate_map_key, st.coordinate_map_key, 3)
            want = ops.conv_table(x, conv.kernel.detach(), kmap.nbr, kmap.n_out, 27, 16, 32, tile_mask=kmap.tile_mask)
            assert torch.equal(y.detach(), want)  # packed image == image packed from the current weights
            y.square().mean().backward()
            before = conv.kernel.detach().clone()
            opt.step()
            opt.zero_grad(set_to_none=True)
            assert not torch.equal(before, conv.kernel.detach())
    with torch.no_grad():
        conv.kernel.mul_(0.5)  # in-place edit: version counter
    st = ME.SparseTensor(features=x, coordinates=coords)
    kmap = st.coordinate_manager.kernel_map(st.coordinate_map_key, st.coordinate_map_key, 3)
    assert torch.equal(conv(st).F.detach(), ops.conv_table(x, conv.kernel.detach(), kmap.nbr, kmap.n_out, 27, 16, 32,
                                                           tile_mask=kmap.tile_mask))


@pytest.mark.parametrize("dtype", [torch.int16, torch.int64])
def test_fused_cross_entropy_matches_torch(dtype):
    """ops.cross_entropy (fused forward / backward kernels) vs F.cross_entropy(ignore_index=-1): loss to 1e-6
    relative (double-precision block sums vs torch's float accumulation), gradient to 1e-6 absolute of its scale."""
    from minsu3d_b200 import ops
    torch.manual_seed(3)
    n, c = 123_457, 20
    x = (torch.randn(n, c, device="cuda") * 3).requires_grad_(True)
    lab = torch.randint(-1, c, (n,), device="cuda").to(dtype)
    want = torch.nn.functional.cross_entropy(x, lab.long(), ignore_index=-1)
    (want * 1.7).backward()
    gw = x.grad.clone()
    x.grad = None
    got = ops.cross_entropy(x, lab, ignore_index=-1)
    (got * 1.7).backward()
    assert abs(float(got) - float(want)) < 1e-6 * abs(float(want))
    assert float((x.grad - gw).abs().max()) < 1e-6 * float(gw.abs().max()) + 1e-12
    assert torch.equal(ops.cross_entropy(x.detach(), lab), ops.cross_entropy(x.detach(), lab))  # deterministic
