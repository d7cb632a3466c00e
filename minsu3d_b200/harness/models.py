"""The reference's callers of the hot path, without the Lightning/Hydra control plane.

These modules exist so that the path can be exercised, parity-tested and benchmarked end to end
(BASELINE.json configs 1-4).  Structure, parameter names and hyper-parameters follow
minsu3d/model/module/{common,backbone,tiny_unet}.py and minsu3d/model/{general_model,pointgroup,
hais,softgroup}.py (state-dict keys match: `backbone.unet.0.kernel`, `...conv_branch.0.bn.weight`),
but every sparse op goes through minsu3d_b200 (ME shim + common_ops mirror) and the clustering
stage stays on the device: no `.cpu()` round trips (pointgroup.py:41-63 has six).
"""
import os
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import MinkowskiEngine as ME
from .. import ops
from ..MinkowskiEngine import modules as me_modules
from ..common_ops.functions import common_ops, hais_ops, pointgroup_ops, softgroup_ops
from . import scenes


# ------------------------------------------------------------------------------------------
# MinkUNet pieces (common.py:21-95, backbone.py:8-43, tiny_unet.py:7-19)
# ------------------------------------------------------------------------------------------
ASYNC_SIZES = True   # use the loader's level sizes / uniqueness guarantees instead of host reads (validated on device)
FUSED_UPDOWN = True   # BN->ReLU->(de)conv of the U-Net levels as one call (b2s_bnconv_*; tests/test_gpu_models.py)
FUSED_BLOCKS = True  # training-mode residual blocks through b2s_resblock_forward/backward (host-side fusion)


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, out_channels, dimension=3, norm_fn=None):
        super().__init__()
        norm_fn = norm_fn or ME.MinkowskiBatchNorm
        self.downsample = None
        if in_channels != out_channels:
            self.downsample = nn.Sequential(
                ME.MinkowskiConvolution(in_channels, out_channels, kernel_size=1, dimension=dimension))
        self.conv_branch = nn.Sequential(
            norm_fn(in_channels), ME.MinkowskiReLU(inplace=True),
            ME.MinkowskiConvolution(in_channels, out_channels, kernel_size=3, dimension=dimension),
            norm_fn(out_channels), ME.MinkowskiReLU(inplace=True),
            ME.MinkowskiConvolution(out_channels, out_channels, kernel_size=3, dimension=dimension))

    def forward(self, x):
        if FUSED_BLOCKS:
            # one autograd node + one library call for the whole block (csrc/fused.cu); bit-identical to the
            # module-by-module path below, which stays the reference-shaped fallback (eval mode, odd shapes)
            br = self.conv_branch
            ds = None if self.downsample is None else self.downsample[0]
            if (isinstance(br[0], ME.MinkowskiBatchNorm) and isinstance(br[3], ME.MinkowskiBatchNorm)
                    and me_modules.residual_block_fusable(x, br[0], br[2], br[3], br[5], ds)):
                return me_modules.fused_residual_block(x, br[0], br[2], br[3], br[5], ds)
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.conv_branch(x)
        y += shortcut
        return y


class UBlock(nn.Module):
    def __init__(self, n_planes, norm_fn, block_reps, block):
        super().__init__()
        self.nPlanes = list(n_planes)
        c = self.nPlanes[0]
        self.blocks = nn.Sequential(OrderedDict(
            ("block%d" % i, block(c, c, 3, norm_fn)) for i in range(block_reps)))
        if len(self.nPlanes) > 1:
            c1 = self.nPlanes[1]
            self.conv = nn.Sequential(norm_fn(c), ME.MinkowskiReLU(inplace=True),
                                      ME.MinkowskiConvolution(c, c1, kernel_size=2, stride=2, dimension=3))
            self.u = UBlock(self.nPlanes[1:], norm_fn, block_reps, block)
            self.deconv = nn.Sequential(norm_fn(c1), ME.MinkowskiReLU(inplace=True),
                                        ME.MinkowskiConvolutionTranspose(c1, c, kernel_size=2, stride=2, dimension=3))
            self.blocks_tail = nn.Sequential(OrderedDict(
                ("block%d" % i, block(c * (2 - i), c, 3, norm_fn)) for i in range(block_reps)))

    @staticmethod
    def _bn_relu_conv(seq, x):
        if FUSED_UPDOWN and me_modules.bn_relu_conv_fusable(x, seq[0], seq[2]):
            return me_modules.fused_bn_relu_conv(x, seq[0], seq[2])
        return seq(x)

    def forward(self, x):
        out = self.blocks(x)
        if len(self.nPlanes) > 1:
            skip = out
            out = self._bn_relu_conv(self.deconv, self.u(self._bn_relu_conv(self.conv, out)))
            out = self.blocks_tail(ME.cat(skip, out))
        return out


EARLY_LOSSES = os.environ.get("B2S_EARLY_LOSSES", "1") != "0"

# what the backbone reads; Trainer.step_from_host copies these first and the rest of the batch on a copy stream
BACKBONE_INPUTS = ("voxel_features", "voxel_xyz", "voxel_point_map", "voxel_level_sizes")


def wait_late_inputs(data):
    """Make the compute stream wait for the part of the batch that Trainer.step_from_host copies on its copy stream."""
    ev = data.get("_late_event")
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)


class PointLinear(nn.Linear):
    """nn.Linear (same parameters / state-dict names) whose weight gradient over many points runs on libb2s."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class Backbone(nn.Module):
    def __init__(self, input_channel, output_channel, block_channels, block_reps, sem_classes):
        super().__init__()
        m = output_channel
        self.unet = nn.Sequential(
            ME.MinkowskiConvolution(input_channel, m, kernel_size=3, dimension=3),
            UBlock([m * c for c in block_channels], ME.MinkowskiBatchNorm, block_reps, ResidualBlock),
            ME.MinkowskiBatchNorm(m), ME.MinkowskiReLU(inplace=True))
        self.semantic_branch = nn.Sequential(PointLinear(m, m), nn.BatchNorm1d(m), nn.ReLU(inplace=True),
                                             PointLinear(m, sem_classes))
        self.offset_branch = nn.Sequential(PointLinear(m, m), nn.BatchNorm1d(m), nn.ReLU(inplace=True),
                                           PointLinear(m, 3))

    def forward(self, voxel_features, voxel_coordinates, v2p_map, level_sizes=None):
        # the loader's voxels are unique by construction (sparse_quantize) and it reports the level sizes: no host
        # read of device counts in the whole backbone, the host keeps enqueuing while the GPU finishes the last step
        # (uniqueness is the loader's contract -- the voxels come out of sparse_quantize, general_dataset.py:159 -- and is
        # validated on the device; level sizes are optional loader metadata, bench.py --size-hints)
        x = ME.SparseTensor(features=voxel_features, coordinates=voxel_coordinates,
                            coordinates_unique=ASYNC_SIZES, level_sizes=level_sizes)
        unet_out = self.unet(x)
        point_features = ops.devoxelize(unet_out.features, v2p_map)  # == features[v2p_map], backbone.py:40
        return {"point_features": point_features,
                "semantic_scores": self.semantic_branch(point_features),
                "point_offsets": self.offset_branch(point_features)}


class TinyUnet(nn.Module):
    def __init__(self, channel):
        super().__init__()
        self.unet = nn.Sequential(UBlock([channel, 2 * channel], ME.MinkowskiBatchNorm, 2, ResidualBlock),
                                  ME.MinkowskiBatchNorm(channel), ME.MinkowskiReLU(inplace=True))

    def forward(self, x):
        return self.unet(x)


# ------------------------------------------------------------------------------------------
# configuration (config/model/{pointgroup,hais,softgroup}.yaml, config/data/scannetv2.yaml)
# ------------------------------------------------------------------------------------------
@dataclass
class Config:
    model: str = "pointgroup"
    m: int = 16
    blocks: List[int] = field(default_factory=lambda: [1, 2, 3, 4, 5, 6, 7])
    block_reps: int = 2
    use_color: bool = True
    use_normal: bool = False
    classes: int = scenes.NUM_CLASSES
    ignore_classes: tuple = scenes.IGNORE_CLASSES
    point_num_avg: List[float] = field(default_factory=lambda: list(scenes.POINT_NUM_AVG))
    radius_avg: List[float] = field(default_factory=lambda: list(scenes.RADIUS_AVG))
    fg_thresh: float = 0.75
    bg_thresh: float = 0.25
    score_scale: int = 50
    score_fullscale: int = 14
    cluster_radius: float = 0.03
    cluster_meanActive: int = 50
    cluster_shift_meanActive: int = 300
    cluster_npoint_thre: int = 50
    # HAIS
    using_set_aggr_in_training: bool = False
    using_set_aggr_in_testing: bool = True
    use_mask_filter_score_feature: bool = False
    mask_filter_score_feature_thre: float = 0.5
    cal_iou_based_on_mask: bool = False
    # SoftGroup
    sg_score_thr: float = 0.2
    sg_radius: float = 0.04
    sg_mean_active: int = 300
    sg_npoint_thr: float = 0.05
    sg_min_npoint: int = 100
    sg_max_proposal_num: int = 200
    sg_pos_iou_thr: float = 0.5
    # where the clustering stage takes its semantic predictions / offsets from:
    #   "network": the network's own outputs (reference behaviour, pointgroup.py:28-47)
    #   "gt_noise": ground truth + noise, so that random-init weights still yield proposals
    proposal_source: str = "network"
    lr: float = 0.002

    @staticmethod
    def for_model(name, **kw):
        base = dict(model=name)
        if name == "hais":
            base.update(m=32, score_fullscale=20, lr=0.0015, fg_thresh=1.0, bg_thresh=0.0)  # hais.yaml:31-35
        elif name == "softgroup":
            base.update(m=32, score_fullscale=20, lr=0.004)
        base.update(kw)
        return Config(**base)


def clusters_voxel_coords_torch(clusters_idx, clusters_offset, coords, scale, spatial_shape, rand):
    """The reference's torch expression sequence (general_model.py:154-184) up to `batched_xyz`: kept as the
    restatement the fused kernel is checked against (tests/test_gpu_cluster_ops.py)."""
    batch_idx = clusters_idx[:, 0]
    cc = coords[clusters_idx[:, 1]].contiguous()
    mean = common_ops.sec_mean(cc, clusters_offset)
    cc = cc - torch.index_select(mean, 0, batch_idx)
    cmin = common_ops.sec_min(cc, clusters_offset)
    cmax = common_ops.sec_max(cc, clusters_offset)
    cscale = 1 / ((cmax - cmin) / spatial_shape).max(1)[0] - 0.01  # ensures voxel coords < spatial_shape
    cscale = torch.clamp(cscale, min=None, max=scale)
    min_xyz, max_xyz = cmin * cscale[:, None], cmax * cscale[:, None]
    cc = cc * torch.index_select(cscale, 0, batch_idx)[:, None]
    rng = max_xyz - min_xyz
    offset = -min_xyz + torch.clamp(spatial_shape - rng - 0.001, min=0) * rand[0]
    offset = offset + torch.clamp(spatial_shape - rng + 0.001, max=0) * rand[1]
    cc = cc + torch.index_select(offset, 0, batch_idx)
    cc = cc.int()
    return torch.cat((clusters_idx[:, 0].unsqueeze(-1).to(cc.dtype), cc), dim=1)


def clusters_voxelization(clusters_idx, clusters_offset, feats, coords, scale, spatial_shape, rand=None):
    """general_model.py:152-193 with the coordinate part fused into one library call.  `rand`
    ([2,3] tensor) replaces the two torch.rand(3) draws so that parity runs can share them (appendix C.11)."""
    device = feats.device
    feats = feats[clusters_idx[:, 1]]
    if rand is None:
        rand = torch.rand(2, 3, device=device)
    batched_xyz = ops.clusters_voxelize(clusters_idx.contiguous(), clusters_offset, coords, scale, spatial_shape, rand)
    voxel_xyz, voxel_features, _, voxel_point_map = ME.utils.sparse_quantize(
        batched_xyz, feats, return_index=True, return_inverse=True, device="cuda")
    # sparse_quantize output is unique by construction: no second host read for the row count
    return ME.SparseTensor(features=voxel_features, coordinates=voxel_xyz, device=device,
                           coordinates_unique=ASYNC_SIZES), voxel_point_map


def get_segmented_scores(scores, fg_thresh=1.0, bg_thresh=0.0):
    """general_model.py:196-213: 1 above fg, 0 below bg, linear in between."""
    k = 1 / (fg_thresh - bg_thresh)
    b = bg_thresh / (bg_thresh - fg_thresh)
    return torch.where(scores > fg_thresh, torch.ones_like(scores),
                       torch.where(scores < bg_thresh, torch.zeros_like(scores), scores * k + b))


def pt_offset_loss(pred_offsets, gt_offsets, valid_mask):
    """minsu3d/loss/pt_offset_loss.py:12-38.  Same means over the valid points, written as masked sums: the reference's
    `pred[valid_mask]` costs two host reads (count_nonzero, the boolean index) and a sort-based index_put in the backward
    (~0.3 ms per step); the masked form has no data-dependent shape.  No valid point -> both losses are 0 like the
    reference's early return."""
    m = valid_mask.to(pred_offsets.dtype)
    cnt = m.sum().clamp(min=1.0)
    norm_loss = (torch.sum(torch.abs(pred_offsets - gt_offsets), dim=-1) * m).sum() / cnt
    eps = torch.finfo(gt_offsets.dtype).eps
    cos = (F.normalize(gt_offsets, p=2, dim=1, eps=eps) * F.normalize(pred_offsets, p=2, dim=1, eps=eps)).sum(-1)
    dir_loss = -(cos * m).sum() / cnt
    return norm_loss, dir_loss


class GeneralModel(nn.Module):
    """general_model.py:16-50 without Lightning: backbone forward + semantic/offset losses."""

    def __init__(self, cfg: Config):
        super().__init__()
        self.cfg = cfg
        in_ch = 3 + 3 * cfg.use_color + 3 * cfg.use_normal
        self.backbone = Backbone(in_ch, cfg.m, cfg.blocks, cfg.block_reps, cfg.classes)
        self.clustering = True  # current_epoch > prepare_epochs

    def backbone_forward(self, data):
        return self.backbone(data["voxel_features"], data["voxel_xyz"], data["voxel_point_map"],
                             data.get("voxel_level_sizes") if ASYNC_SIZES else None)

    def base_loss(self, data, out):
        if "_base_losses" in out:  # already enqueued while the host waited for the clustering counts (forward)
            return dict(out["_base_losses"])
        if out["semantic_scores"].is_cuda:  # fused libb2s kernels (torch's nll_loss reduction is one CTA: 0.67 ms)
            losses = {"semantic_loss": ops.cross_entropy(out["semantic_scores"], data["sem_labels"], ignore_index=-1)}
        else:
            losses = {"semantic_loss": F.cross_entropy(out["semantic_scores"], data["sem_labels"].long(), ignore_index=-1)}
        gt_offsets = data["instance_center_xyz"] - data["point_xyz"]
        losses["offset_norm_loss"], losses["offset_dir_loss"] = pt_offset_loss(
            out["point_offsets"], gt_offsets, data["instance_ids"] != -1)
        return losses

    # semantic predictions / offsets that drive the clustering stage
    def _cluster_inputs(self, data, out):
        if self.cfg.proposal_source == "gt_noise":
            g = torch.Generator(device=out["semantic_scores"].device)
            g.manual_seed(1234)
            sem = data["sem_labels"].long().clamp(min=0)
            scores = F.one_hot(sem, self.cfg.classes).float() * 8 + torch.randn(
                out["semantic_scores"].shape, device=sem.device, generator=g)
            shrink = 0.7 + 0.3 * torch.rand((sem.numel(), 1), device=sem.device, generator=g)
            offsets = (data["instance_center_xyz"] - data["point_xyz"]) * shrink
            offsets = torch.where((data["instance_ids"] != -1)[:, None], offsets, torch.zeros_like(offsets))
            return scores, offsets
        return out["semantic_scores"], out["point_offsets"]

    def _object_points(self, semantic_preds):
        mask = torch.ones_like(semantic_preds, dtype=torch.bool)
        for c in self.cfg.ignore_classes:
            mask &= semantic_preds != (c - 1)
        return torch.nonzero(mask).view(-1)

    def training_loss(self, data):
        out = self(data)
        losses = self.loss(data, out)
        return sum(losses.values()), losses, out


class PointGroup(GeneralModel):
    """pointgroup.py:12-109."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.score_net = TinyUnet(cfg.m)
        self.score_branch = nn.Linear(cfg.m, 1)

    def forward(self, data, rand=None):
        cfg = self.cfg
        out = self.backbone_forward(data)
        wait_late_inputs(data)
        if not self.clustering:
            return out
        scores, offsets = self._cluster_inputs(data, out)
        semantic_preds = scores.argmax(1).to(torch.int16)
        object_idxs = self._object_points(semantic_preds)
        batch_idxs_ = data["vert_batch_ids"][object_idxs]
        batch_offsets_ = torch.cumsum(torch.bincount(batch_idxs_ + 1), dim=0).int()
        coords_ = data["point_xyz"][object_idxs]
        sem_ = semantic_preds[object_idxs].contiguous()
        # both ball queries, then both BFS clusterings, each pair with ONE host read of its data-dependent sizes
        # (the reference-shaped single calls -- common_ops.ballquery_batch_p / pointgroup_ops.pg_bfs_cluster -- read
        # one count each; results are identical, tests/test_gpu_models.py)
        grad_on = torch.is_grad_enabled()
        with torch.no_grad():
            shifted = (coords_ + offsets.detach()[object_idxs]).contiguous()
            # the semantic / offset losses do not depend on the proposals: they are enqueued between the launch of the
            # ball-query count passes and the host read of their results, so their ~1 ms of host work (60 small
            # kernels) hides behind the GPU instead of delaying the backward
            def early_losses():
                if grad_on and self.training and EARLY_LOSSES:
                    with torch.enable_grad():
                        out["_base_losses"] = self.base_loss(data, out)
            queries = ops.ballquery_many([coords_.contiguous(), shifted], batch_idxs_.contiguous(), batch_offsets_,
                                         cfg.cluster_radius, between=early_losses)
            sets = []
            for p_idx, p_off in ops.pg_cluster_many(sem_, queries, cfg.cluster_npoint_thre):
                p_idx = p_idx.long()
                p_idx[:, 1] = object_idxs[p_idx[:, 1]]
                sets.append((p_idx, p_off))
        (p_idx, p_off), (s_idx, s_off) = sets  # unshifted first, shifted appended (pointgroup.py:70-73)
        s_idx[:, 0] += p_off.size(0) - 1
        proposals_idx = torch.cat((p_idx, s_idx), dim=0)
        proposals_offset = torch.cat((p_off, s_off[1:] + p_off[-1]))
        out["proposal_scores"] = None
        if proposals_offset.numel() > 1:
            vox, p2v = clusters_voxelization(proposals_idx, proposals_offset, out["point_features"], data["point_xyz"],
                                             cfg.score_scale, cfg.score_fullscale, rand)
            score_feats = self.score_net(vox)
            pt_score_feats = ops.devoxelize(score_feats.features, p2v)
            proposals_score_feats = common_ops.roipool(pt_score_feats, proposals_offset)
            out["proposal_scores"] = (self.score_branch(proposals_score_feats), proposals_idx, proposals_offset)
        return out

    def loss(self, data, out):
        losses = self.base_loss(data, out)
        if self.clustering and out.get("proposal_scores") is not None:
            scores, proposals_idx, proposals_offset = out["proposal_scores"]
            ious = common_ops.get_iou(proposals_idx[:, 1].int().contiguous(), proposals_offset,
                                      data["instance_ids"], data["instance_num_point"])
            gt_scores = get_segmented_scores(ious.max(1)[0], self.cfg.fg_thresh, self.cfg.bg_thresh)
            losses["score_loss"] = F.binary_cross_entropy_with_logits(scores.view(-1), gt_scores)
        return losses


class HAIS(GeneralModel):
    """hais.py:12-127."""

    def __init__(self, cfg):
        super().__init__(cfg)
        m = cfg.m
        self.tiny_unet = TinyUnet(m)
        self.score_branch = nn.Linear(m, 1)
        self.mask_branch = nn.Sequential(PointLinear(m, m), nn.ReLU(inplace=True), PointLinear(m, 1))

    def forward(self, data, rand=None):
        cfg = self.cfg
        out = self.backbone_forward(data)
        wait_late_inputs(data)
        if not self.clustering:
            return out
        scores, offsets = self._cluster_inputs(data, out)
        semantic_preds = scores.argmax(1).to(torch.int16)
        object_idxs = self._object_points(semantic_preds)
        batch_idxs_ = data["vert_batch_ids"][object_idxs]
        batch_offsets_ = torch.cumsum(torch.bincount(batch_idxs_ + 1), dim=0).int()
        offset_coords_ = (data["point_xyz"][object_idxs] + offsets.detach()[object_idxs]).contiguous()
        idx, start_len = common_ops.ballquery_batch_p(offset_coords_, batch_idxs_, batch_offsets_, cfg.cluster_radius,
                                                      cfg.cluster_shift_meanActive)
        set_aggr = cfg.using_set_aggr_in_training if self.training else cfg.using_set_aggr_in_testing
        proposals_idx, proposals_offset = hais_ops.hierarchical_aggregation(
            semantic_preds[object_idxs].contiguous(), offset_coords_, idx, start_len, batch_idxs_.contiguous(),
            set_aggr, cfg.point_num_avg, cfg.radius_avg, -1)
        proposals_idx = proposals_idx.long()
        proposals_idx[:, 1] = object_idxs[proposals_idx[:, 1]]
        out["proposal_scores"] = None
        if proposals_offset.numel() > 1:
            vox, p2v = clusters_voxelization(proposals_idx, proposals_offset, out["point_features"], data["point_xyz"],
                                             cfg.score_scale, cfg.score_fullscale, rand)
            inst = self.tiny_unet(vox)
            score_feats = ops.devoxelize(inst.features, p2v)
            mask_scores = self.mask_branch(inst.features)[p2v]  # linear first: fewer voxels than points
            if cfg.use_mask_filter_score_feature:
                keep = (torch.sigmoid(mask_scores) >= cfg.mask_filter_score_feature_thre).float()
                score_feats = score_feats * keep
            score_feats = common_ops.roipool(score_feats.contiguous(), proposals_offset)
            out["proposal_scores"] = (self.score_branch(score_feats), proposals_idx, proposals_offset, mask_scores)
        return out

    def loss(self, data, out):
        losses = self.base_loss(data, out)
        if self.clustering and out.get("proposal_scores") is not None:
            scores, proposals_idx, proposals_offset, mask_scores = out["proposal_scores"]
            sig = torch.sigmoid(mask_scores)
            pidx = proposals_idx[:, 1].int().contiguous()
            if self.cfg.cal_iou_based_on_mask:
                ious = common_ops.get_mask_iou_on_pred(pidx, proposals_offset, data["instance_ids"],
                                                       data["instance_num_point"], sig.detach().contiguous())
            else:
                ious = common_ops.get_mask_iou_on_cluster(pidx, proposals_offset, data["instance_ids"],
                                                          data["instance_num_point"])
            mask_label, mask_label_mask = common_ops.get_mask_label(
                pidx, proposals_offset, data["instance_ids"], data["instance_semantic_cls"],
                data["instance_num_point"], ious, -1, 0.5)
            losses["mask_loss"] = F.binary_cross_entropy(sig, mask_label.unsqueeze(1).float(),
                                                         weight=mask_label_mask.unsqueeze(1).float(), reduction="mean")
            gt_scores = get_segmented_scores(ious.max(1)[0], self.cfg.fg_thresh, self.cfg.bg_thresh)
            losses["score_loss"] = F.binary_cross_entropy_with_logits(scores.view(-1), gt_scores)
        return losses


def soft_grouping_loop(cfg, semantic_scores, offsets, point_xyz, vert_batch_ids):
    """softgroup.py:43-86 as written: one ball query + BFS clustering per semantic class (kept as the restatement the
    batched version below is checked against)."""
    idx_list, off_list = [], []
    n_prop, n_pts = 0, 0
    for class_id in range(cfg.classes):
        if class_id + 1 in cfg.ignore_classes:
            continue
        object_idxs = (semantic_scores[:, class_id] > cfg.sg_score_thr).nonzero().view(-1)
        if object_idxs.size(0) < cfg.sg_min_npoint:
            continue
        batch_idxs_ = vert_batch_ids[object_idxs]
        batch_offsets_ = torch.cumsum(torch.bincount(batch_idxs_ + 1), dim=0).int()
        xyz = (point_xyz[object_idxs] + offsets[object_idxs]).contiguous()
        idx, start_len = common_ops.ballquery_batch_p(xyz, batch_idxs_, batch_offsets_, cfg.sg_radius, cfg.sg_mean_active)
        p_idx, p_off = softgroup_ops.sg_bfs_cluster(cfg.point_num_avg, idx, start_len, cfg.sg_npoint_thr, class_id)
        if p_idx.size(0) == 0:
            continue
        p_idx = p_idx.long()
        p_idx[:, 1] = object_idxs[p_idx[:, 1]]
        p_idx[:, 0] += n_prop
        idx_list.append(p_idx)
        off_list.append(p_off[1:] + n_pts if off_list else p_off)
        n_prop += p_off.numel() - 1
        n_pts += p_idx.size(0)
    if not idx_list:
        return None
    return torch.cat(idx_list, dim=0), torch.cat(off_list)


def soft_grouping(cfg, semantic_scores, offsets, point_xyz, vert_batch_ids):
    """All classes of softgroup.py:43-86 in one ball query and one clustering pass: the per-class point sets are
    stacked class by class and every (class, scene) pair gets its own batch index, so neighbour lists never cross
    classes; cluster_select mode 3 applies the per-class size threshold.  Proposals come out in the reference's
    order (classes ascending, seeds ascending inside a class).  Two host reads instead of ~4 per class."""
    n_batch = int(vert_batch_ids.max().item()) + 1 if vert_batch_ids.numel() else 1
    classes = [c for c in range(cfg.classes) if c + 1 not in cfg.ignore_classes]
    if not classes or len(classes) * n_batch > 255:  # composite batch index is uint8 like the reference's
        return soft_grouping_loop(cfg, semantic_scores, offsets, point_xyz, vert_batch_ids)
    dev = semantic_scores.device
    cls_t = torch.tensor(classes, device=dev)
    mask = (semantic_scores[:, cls_t] > cfg.sg_score_thr).t()                 # [n_cls, N]
    mask = mask & (mask.sum(1, keepdim=True) >= cfg.sg_min_npoint)            # classes below min_npoint are skipped
    slot, pts = mask.nonzero(as_tuple=True)                                   # class-major, points ascending
    if pts.numel() == 0:
        return None
    class_id = cls_t[slot]
    batch_idxs_ = (slot * n_batch + vert_batch_ids[pts].long()).to(torch.uint8)
    batch_offsets_ = torch.cumsum(torch.bincount(batch_idxs_.long() + 1, minlength=len(classes) * n_batch + 1),
                                  dim=0).int()
    xyz = (point_xyz[pts] + offsets[pts]).contiguous()
    idx, start_len = ops.ballquery(xyz, batch_idxs_.contiguous(), batch_offsets_, cfg.sg_radius)
    mean = torch.tensor(cfg.point_num_avg, dtype=torch.float32, device=dev)
    thr = torch.full_like(mean, float(cfg.sg_npoint_thr))
    thr = torch.where(mean != -1, thr * mean, thr)                             # bfs_cluster.cpp:116-121, fp32
    comp = ops.cluster_label(idx, start_len, None)
    p_idx, p_off = ops.cluster_extract(idx, start_len, class_id.to(torch.int16).contiguous(), comp, mode=3,
                                       point_num_avg=thr.contiguous())
    if p_idx.size(0) == 0:
        return None
    p_idx = p_idx.long()
    p_idx[:, 1] = pts[p_idx[:, 1]]
    return p_idx, p_off


class SoftGroup(GeneralModel):
    """softgroup.py:11-183."""

    def __init__(self, cfg):
        super().__init__(cfg)
        m = cfg.m
        self.instance_classes = cfg.classes - len(cfg.ignore_classes)
        self.tiny_unet = TinyUnet(m)
        self.classification_branch = nn.Linear(m, self.instance_classes + 1)
        self.mask_scoring_branch = nn.Sequential(nn.Linear(m, m), nn.ReLU(inplace=True),
                                                 nn.Linear(m, self.instance_classes + 1))
        self.iou_score = nn.Linear(m, self.instance_classes + 1)

    def forward(self, data, rand=None):
        cfg = self.cfg
        out = self.backbone_forward(data)
        wait_late_inputs(data)
        if not self.clustering:
            return out
        scores, offsets = self._cluster_inputs(data, out)
        semantic_scores = scores.softmax(dim=-1)
        offsets = offsets.detach()
        proposals = soft_grouping(cfg, semantic_scores, offsets, data["point_xyz"], data["vert_batch_ids"])
        idx_list, off_list = ([proposals[0]], [proposals[1]]) if proposals is not None else ([], [])
        out["proposals_idx"] = None
        if not idx_list:
            return out
        proposals_idx = torch.cat(idx_list, dim=0)
        proposals_offset = torch.cat(off_list)
        if proposals_offset.shape[0] > cfg.sg_max_proposal_num:
            proposals_offset = proposals_offset[:cfg.sg_max_proposal_num + 1]
            proposals_idx = proposals_idx[:int(proposals_offset[-1])]
        out["proposals_idx"], out["proposals_offset"] = proposals_idx, proposals_offset
        vox, inst_map = clusters_voxelization(proposals_idx, proposals_offset, out["point_features"], data["point_xyz"],
                                              cfg.score_scale, cfg.score_fullscale, rand)
        feats = self.tiny_unet(vox)
        mask_scores = self.mask_scoring_branch(feats.features)
        out["mask_scores"] = mask_scores[inst_map]
        out["instance_batch_idxs"] = feats.coordinates[:, 0][inst_map]
        # global_pool (softgroup.py:112-115): voxel rows are grouped by proposal because the
        # clusters arrive proposal by proposal and sparse_quantize keeps first-occurrence order
        indices = feats.coordinates[:, 0]
        batch_offset = torch.cumsum(torch.bincount(indices + 1), dim=0).int()
        pooled = softgroup_ops.global_avg_pool(feats.features.contiguous(), batch_offset)
        out["cls_scores"] = self.classification_branch(pooled)
        out["iou_scores"] = self.iou_score(pooled)
        return out

    def loss(self, data, out):
        losses = self.base_loss(data, out)
        if not (self.clustering and out.get("proposals_idx") is not None):
            return losses
        cfg = self.cfg
        pidx = out["proposals_idx"][:, 1].int().contiguous()
        poff = out["proposals_offset"]
        ious_on_cluster = common_ops.get_mask_iou_on_cluster(pidx, poff, data["instance_ids"], data["instance_num_point"])
        fg_inds = data["instance_semantic_cls"] != -1
        fg_instance_cls = data["instance_semantic_cls"][fg_inds]
        fg_ious = ious_on_cluster[:, fg_inds]
        n_prop = fg_ious.size(0)
        assigned = fg_ious.new_full((n_prop,), -1, dtype=torch.long)
        max_iou, argmax_iou = fg_ious.max(1)
        pos = max_iou >= cfg.sg_pos_iou_thr
        assigned[pos] = argmax_iou[pos]
        labels = fg_instance_cls.new_full((n_prop,), self.instance_classes)
        pos = assigned >= 0
        labels[pos] = fg_instance_cls[assigned[pos]]
        labels = labels.long()
        losses["classification_loss"] = F.cross_entropy(out["cls_scores"], labels)
        mask_cls_label = labels[out["instance_batch_idxs"].long()]
        rows = torch.arange(mask_cls_label.size(0), device=labels.device)
        mask_sig = out["mask_scores"].sigmoid()[rows, mask_cls_label]
        mask_label, mask_label_mask = common_ops.get_mask_label(
            pidx, poff, data["instance_ids"], data["instance_semantic_cls"], data["instance_num_point"],
            ious_on_cluster, -1, cfg.sg_pos_iou_thr)
        msl = F.binary_cross_entropy(mask_sig, mask_label.float(), weight=mask_label_mask.float(), reduction="sum")
        losses["mask_scoring_loss"] = msl / (torch.count_nonzero(mask_label_mask) + 1)
        ious = common_ops.get_mask_iou_on_pred(pidx, poff, data["instance_ids"], data["instance_num_point"],
                                               mask_sig.detach().contiguous())
        rows = torch.arange(labels.size(0), device=labels.device)
        weight = labels < self.instance_classes
        iou_slice = out["iou_scores"][rows, labels]
        iou_loss = F.mse_loss(iou_slice, ious[:, fg_inds].max(1)[0], reduction="none")
        losses["iou_scoring_loss"] = iou_loss[weight].sum() / (weight.count_nonzero() + 1)
        return losses


MODELS = {"pointgroup": PointGroup, "hais": HAIS, "softgroup": SoftGroup}


def build_model(cfg: Config):
    return MODELS[cfg.model](cfg)
