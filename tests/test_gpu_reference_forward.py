"""The reference's OWN model code (minsu3d/model/{general_model,pointgroup,hais,softgroup}.py, module/*.py,
common_ops/functions/*.py, unmodified, staged by oracle/build_ref.py) executed on the GPU on top of the drop-in
(`import MinkowskiEngine` / `import COMMON_OPS` -> minsu3d_b200), compared with the harness models that bench.py
times: same weights, same batch, shared torch.rand draws.  This is the proof of boundaries #1 and #2
(pointgroup.py:23-109, backbone.py:36-43, general_model.py:152-193)."""
import numpy as np
import pytest
import torch

from ref_loader import load_reference_models

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    r = load_reference_models()
    if r is None:
        pytest.skip("reference python not staged (oracle/build_ref.py needs /root/reference)")
    return r


@pytest.fixture(scope="module")
def batch():
    """Two 20k-point scenes whose colour is a function of the semantic class (+ noise): the reference's clustering
    stage is driven by the NETWORK's predictions (pointgroup.py:28-47), and it produces proposals only if the network
    has learnt the semantics -- which a short training run can do when the class is visible in the features."""
    from minsu3d_b200.harness import scenes
    rng = np.random.default_rng(5)
    palette = rng.uniform(-1, 1, (scenes.NUM_CLASSES, 3)).astype(np.float32)
    out = []
    for seed in (21, 22):
        sc = scenes.make_scene(seed, 20_000)
        sc["rgb"] = (palette[sc["sem_labels"]] + rng.normal(0, 0.05, sc["rgb"].shape)).astype(np.float32)
        out.append(sc)
    return scenes.collate(out, "cuda")


def _pretrain(name, batch, steps=250):
    """Random-init weights predict noise -> no proposals (SURVEY 8 caveat).  A short run of the harness trainer
    on the fixed batch (semantic + offset losses only) gives weights whose predictions cluster."""
    from minsu3d_b200.harness import models, train
    cfg = models.Config.for_model(name, proposal_source="network", lr=0.01)
    tr = train.Trainer(cfg, "cuda", seed=7)
    tr.model.clustering = False
    for _ in range(steps):
        tr.step(batch)
    tr.model.clustering = True
    with torch.no_grad():
        acc = float((tr.model.backbone_forward(batch)["semantic_scores"].argmax(1) == batch["sem_labels"]).float().mean())
    assert acc > 0.85, "pretraining did not learn the semantics (accuracy %.2f): no proposals would be found" % acc
    return cfg, tr.model


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _shared_rand(seed):
    torch.manual_seed(seed)
    r = torch.stack((torch.rand(3, device="cuda"), torch.rand(3, device="cuda")))
    torch.manual_seed(seed)  # the reference now draws the same two torch.rand(3)
    return r


@pytest.mark.parametrize("name", ["pointgroup", "hais", "softgroup"])
def test_reference_forward_and_loss_on_dropin_equal_harness(ref, batch, name):
    ref_models, make_cfg = ref
    cfg, own = _pretrain(name, batch)
    if name == "hais":  # the stub's current_epoch is past both start epochs (hais.yaml:42-44)
        cfg.use_mask_filter_score_feature = True
        cfg.cal_iou_based_on_mask = True
    cls = {"pointgroup": ref_models.PointGroup, "hais": ref_models.HAIS, "softgroup": ref_models.SoftGroup}[name]
    theirs = cls(make_cfg(name)).cuda()
    missing, unexpected = theirs.load_state_dict(own.state_dict(), strict=True)
    assert not missing and not unexpected
    own.train()
    theirs.train()
    state = {k: v.clone() for k, v in own.state_dict().items()}

    # harness (what bench.py times): fused residual blocks, device-resident clustering
    own.zero_grad(set_to_none=True)
    rand = _shared_rand(99)
    out_a = own(batch, rand=rand)
    losses_a = own.loss(batch, out_a)
    sum(losses_a.values()).backward()
    grads_a = {k: p.grad.clone() for k, p in own.named_parameters() if p.grad is not None}

    # reference code, module by module, CPU tensors into bfs_cluster like pointgroup.py:41-63
    theirs.load_state_dict(state)  # undo the running-statistics update of the first forward
    _shared_rand(99)
    out_b = theirs(batch)
    losses_b = theirs._loss(batch, out_b)
    sum(losses_b.values()).backward()
    grads_b = {k: p.grad for k, p in theirs.named_parameters() if p.grad is not None}

    for k in ("point_features", "semantic_scores", "point_offsets"):
        assert torch.equal(out_a[k], out_b[k]), k  # same kernels in the same order: bit-identical
    if name == "softgroup":
        assert out_b["proposals_idx"] is not None and out_b["proposals_offset"].numel() > 2, "no proposals"
        assert torch.equal(out_a["proposals_offset"], out_b["proposals_offset"].to(out_a["proposals_offset"].device))
        assert torch.equal(out_a["proposals_idx"], out_b["proposals_idx"].to(out_a["proposals_idx"].device))
        for k in ("cls_scores", "iou_scores", "mask_scores"):
            assert _rel(out_a[k], out_b[k]) < 1e-5, k
    else:
        pa, pb = out_a["proposal_scores"], out_b["proposal_scores"]
        assert pb[2].numel() > 2, "the reference forward produced no proposals"
        assert torch.equal(pa[2], pb[2]) and torch.equal(pa[1], pb[1])  # offsets, (proposal, point) pairs: bit-exact
        assert _rel(pa[0], pb[0]) < 1e-5
        if name == "hais":
            assert _rel(pa[3], pb[3]) < 1e-5
    assert set(losses_a) == set(losses_b)
    for k in losses_a:
        assert abs(float(losses_a[k]) - float(losses_b[k])) <= 1e-5 * max(1.0, abs(float(losses_b[k]))), k
    assert set(grads_a) == set(grads_b)
    for k in grads_a:
        if k.endswith("_branch.0.bias"):  # a bias in front of a BatchNorm: the true gradient is exactly zero (noise)
            continue
        err = float((grads_a[k] - grads_b[k]).norm() / grads_b[k].norm().clamp_min(1e-12))
        assert err < 1e-4, "%s: rel l2 %.3e" % (k, err)
