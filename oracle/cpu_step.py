"""CPU restatement of ONE PointGroup train step (configs[1]) -- the reference arm of bench.py.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  What is timed here is what the reference does in
`training_step` (minsu3d/model/general_model.py:52-66, pointgroup.py:23-109) on its CPU-capable parts:

  backbone forward + backward   oracle/me_unet_grad.py: restated MinkowskiEngine CPU-backend algorithm (hash kernel maps,
                                per-offset gather-GEMM-scatter, OpenMP) -- MinkowskiEngine itself is an un-vendored
                                dependency and cannot be installed, PARITY UNPINNED against its binary
  ball query x2                 oracle.c restatement of the reference's brute-force kernel (bfs_cluster.cu:15-60), OpenMP
  BFS clustering x2             the REFERENCE'S OWN compiled `pg_bfs_cluster` (oracle/_ref, bfs_cluster.cpp:86-166, one
                                thread: the reference code has no parallelism) when it is built, else the oracle port
  clusters_voxelization         oracle.c restatement of general_model.py:152-182 + first-occurrence unique (numpy)
  ScoreNet forward + backward   same restated sparse ops (tiny_unet.py:7-19), roipool / get_iou from the oracle
  losses, Adam                  torch CPU (general_model.py:36-50, pointgroup.py:95-109, pointgroup.yaml:16-18)

The clustering stage is driven like the GPU arm's (`proposal_source="gt_noise"`: ground truth + noise, because
random-init weights yield no proposals -- SURVEY.md section 8 caveat).
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import ballquery, clusters_voxelize, get_iou, pg_bfs_cluster, roipool_fp
from . import build_ref
from .me_unet import _Maps
from .me_unet_grad import _bn_relu, _ublock, backbone_forward


class _RoiPool(torch.autograd.Function):
    """Segmented max (first max wins) with the gradient routed to the arg-max row (roipool.cu:12-57)."""

    @staticmethod
    def forward(ctx, feats, offsets):
        out, maxidx = roipool_fp(feats.detach().numpy(), offsets)
        ctx.maxidx = torch.from_numpy(maxidx.astype(np.int64))
        ctx.n = feats.shape[0]
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, g):
        gin = torch.zeros((ctx.n, g.shape[1]), dtype=g.dtype)
        gin.scatter_add_(0, ctx.maxidx, g)
        return gin, None


class CpuPointGroupStep:
    """model: a harness PointGroup on the CPU (parameter container with the reference's names + the gt_noise helper)."""

    def __init__(self, model, lr=2e-3):
        self.model = model
        self.params = {k: p for k, p in model.named_parameters()}
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)
        self.ref_ops = build_ref.load()  # the reference's own compiled COMMON_OPS (CPU BFS), or None
        self.timing = {}

    def _bfs(self, labels, idx, start_len, thr):
        if self.ref_ops is not None:
            ci = torch.zeros(0, dtype=torch.int32)
            co = torch.zeros(0, dtype=torch.int32)
            self.ref_ops.pg_bfs_cluster(torch.from_numpy(labels), torch.from_numpy(idx), torch.from_numpy(start_len),
                                        ci, co, labels.shape[0], int(thr))
            return ci.numpy(), co.numpy()
        return pg_bfs_cluster(labels, idx, start_len, thr)

    def step(self, data):
        """data: collated CPU batch (harness.scenes.collate(..., 'cpu')).  Returns (loss, seconds per stage)."""
        cfg = self.model.cfg
        t = {}
        t0 = time.perf_counter()
        sd = dict(self.model.state_dict())
        sd.update(self.params)  # leaves that require grad
        out = backbone_forward(sd, data["voxel_features"].numpy(), data["voxel_xyz"].numpy(),
                               data["voxel_point_map"].numpy(), depth=len(cfg.blocks), reps=cfg.block_reps)
        t["backbone_fwd"] = time.perf_counter() - t0
        losses = self.model.base_loss(data, out)
        # ---- clustering stage (pointgroup.py:28-73) ------------------------------------------------------
        t0 = time.perf_counter()
        scores, offsets = self.model._cluster_inputs(data, out)
        preds = scores.argmax(1).to(torch.int16)
        obj = self.model._object_points(preds).numpy()
        bidx = data["vert_batch_ids"].numpy()[obj]
        offs = np.concatenate(([0], np.cumsum(np.bincount(bidx, minlength=1)))).astype(np.int32)
        lab = np.ascontiguousarray(preds.numpy()[obj])
        xyz = data["point_xyz"].numpy()
        sets = []
        t_bq = t_bfs = 0.0
        for pts in (xyz[obj], (xyz + offsets.detach().numpy())[obj]):
            t1 = time.perf_counter()
            idx, sl = ballquery(np.ascontiguousarray(pts), bidx, offs, cfg.cluster_radius)
            t2 = time.perf_counter()
            ci, co = self._bfs(lab, idx, sl, cfg.cluster_npoint_thre)
            t3 = time.perf_counter()
            t_bq += t2 - t1
            t_bfs += t3 - t2
            sets.append((ci.astype(np.int64), co.astype(np.int32)))
        t["ballquery"], t["bfs_cluster"] = t_bq, t_bfs
        (p_idx, p_off), (s_idx, s_off) = sets
        p_idx[:, 1] = obj[p_idx[:, 1]]
        s_idx[:, 1] = obj[s_idx[:, 1]]
        s_idx[:, 0] += p_off.shape[0] - 1
        prop_idx = np.concatenate((p_idx, s_idx))
        prop_off = np.concatenate((p_off, s_off[1:] + p_off[-1])).astype(np.int32)
        t["cluster_stage"] = time.perf_counter() - t0
        # ---- ScoreNet (pointgroup.py:75-93) ---------------------------------------------------------------
        t0 = time.perf_counter()
        if prop_off.shape[0] > 1:
            rand = torch.rand(2, 3).numpy()
            vox = clusters_voxelize(prop_idx, prop_off, xyz, cfg.score_scale, cfg.score_fullscale, rand)
            _, first, inv = np.unique(vox, axis=0, return_index=True, return_inverse=True)
            order = np.argsort(first, kind="stable")
            rank = np.empty_like(order)
            rank[order] = np.arange(order.size)
            uniq, p2v = first[order], rank[inv.reshape(-1)]
            feats = out["point_features"][torch.from_numpy(prop_idx[:, 1])][torch.from_numpy(uniq)]
            maps = _Maps(np.ascontiguousarray(vox[uniq], np.int32))
            y = _ublock(feats, sd, "score_net.unet.0", maps, 1, 2, 2)
            y = _bn_relu(y, sd, "score_net.unet.1.bn")
            pt = y[torch.from_numpy(p2v)]
            pooled = _RoiPool.apply(pt, prop_off)
            score = F.linear(pooled, sd["score_branch.weight"], sd["score_branch.bias"])
            ious = get_iou(np.ascontiguousarray(prop_idx[:, 1].astype(np.int32)), prop_off, data["instance_ids"].numpy(),
                           data["instance_num_point"].numpy())
            from minsu3d_b200.harness.models import get_segmented_scores
            gt = get_segmented_scores(torch.from_numpy(ious).max(1)[0], cfg.fg_thresh, cfg.bg_thresh)
            losses["score_loss"] = F.binary_cross_entropy_with_logits(score.view(-1), gt)
        t["scorenet_fwd"] = time.perf_counter() - t0
        # ---- backward + Adam ----------------------------------------------------------------------------------
        t0 = time.perf_counter()
        total = sum(losses.values())
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        self.opt.step()
        t["backward_adam"] = time.perf_counter() - t0
        self.timing = t
        return float(total.detach()), t
