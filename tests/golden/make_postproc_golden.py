"""Generates tests/golden/postproc_ref.npz by RUNNING the reference's own post-processing methods.

SURVEY 8(f) #2: `PointGroup._get_nms_instances` / `_get_pred_instances` (minsu3d/model/pointgroup.py:197-265),
`HAIS._get_pred_instances` (minsu3d/model/hais.py:210-247) and `SoftGroup._get_pred_instances`
(minsu3d/model/softgroup.py:269-313).  The model modules import Lightning (not installed), so
the two classes cannot be imported; instead the method sources are pulled out of the reference files with `ast` at
generation time and executed unchanged against a stub `self` (nothing of the reference is copied into this repo).

    python tests/golden/make_postproc_golden.py         # needs /root/reference; CPU only

Inputs are seeded and stored next to the outputs, so the fixture is self-contained.
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import sg_mask_scores  # noqa: E402

REF = os.environ.get("MINSU3D_REFERENCE", "/root/reference")


def reference_methods(rel_path, class_name, names):
    """{name: function} compiled from the unmodified source text of the reference's methods."""
    sys.path.insert(0, REF)
    from minsu3d.evaluation.instance_segmentation import rle_decode, rle_encode  # the reference's own helpers
    path = os.path.join(REF, rel_path)
    src = open(path).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    ns = {"np": np, "torch": torch, "rle_encode": rle_encode}
    out = {}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in names:
            code = textwrap.dedent(ast.get_source_segment(src, fn))
            exec(compile(code, path, "exec"), ns)
            out[fn.name] = ns[fn.name]
    out["rle_decode"] = rle_decode
    return out


def make_proposals(rng, n_points, n_obj, per_obj, dup_frac=0.0):
    """Overlapping proposals like PointGroup's two cluster sets: jittered subsets of ground-truth objects.
    Returns proposals_idx int32 [S,2] sorted by proposal id, n_proposals."""
    owner = rng.integers(0, n_obj, n_points)
    rows = []
    p = 0
    for o in range(n_obj):
        pts = np.nonzero(owner == o)[0]
        for _ in range(per_obj):
            frac = rng.uniform(0.3, 1.0)
            sel = pts[rng.random(pts.size) < frac]
            extra = rng.integers(0, n_points, int(rng.integers(0, 30)))  # a few foreign points
            sel = np.unique(np.concatenate((sel, extra)))
            sel = sel[rng.permutation(sel.size)]  # BFS order is not sorted
            rows.append(np.stack((np.full(sel.size, p), sel), 1))
            p += 1
    # some tiny proposals that the npoint threshold removes
    for _ in range(5):
        sel = rng.choice(n_points, int(rng.integers(1, 60)), replace=False)
        rows.append(np.stack((np.full(sel.size, p), sel), 1))
        p += 1
    return np.concatenate(rows).astype(np.int32), p


def hparams(**test):
    ns = types.SimpleNamespace
    return ns(cfg=ns(model=ns(network=ns(test=ns(**test)))))


def run_pointgroup(case, m):
    self = types.SimpleNamespace(hparams=hparams(TEST_NMS_THRESH=case["nms_thr"], TEST_SCORE_THRESH=case["score_thr"],
                                                 TEST_NPOINT_THRESH=case["npoint_thr"]))
    self._get_nms_instances = types.MethodType(m["_get_nms_instances"], self)
    inst = m["_get_pred_instances"](self, "scene0000_00", case["xyz"], torch.from_numpy(case["scores"]),
                                    torch.from_numpy(case["proposals_idx"]).long(), case["n_proposals"],
                                    torch.from_numpy(case["semantic_scores"]), case["num_ignored"])
    return inst


def run_hais(case, m):
    self = types.SimpleNamespace(hparams=hparams(test_mask_score_thre=case["mask_thr"],
                                                 TEST_SCORE_THRESH=case["score_thr"],
                                                 TEST_NPOINT_THRESH=case["npoint_thr"]))
    return m["_get_pred_instances"](self, "scene0000_00", case["xyz"], torch.from_numpy(case["scores"]),
                                    torch.from_numpy(case["proposals_idx"]).long(), case["n_proposals"],
                                    torch.from_numpy(case["mask_scores"]), torch.from_numpy(case["semantic_scores"]),
                                    case["num_ignored"])


def run_softgroup(case, m):
    ns = types.SimpleNamespace
    self = ns(instance_classes=case["instance_classes"],
              hparams=ns(cfg=ns(model=ns(network=ns(test_cfg=ns(mask_score_thr=case["mask_thr"],
                                                                cls_score_thr=case["cls_thr"],
                                                                min_npoint=case["npoint_thr"]))))))
    return m["_get_pred_instances"](self, "scene0000_00", case["xyz"], torch.from_numpy(case["proposals_idx"]).long(),
                                    case["xyz"].shape[0], torch.from_numpy(case["sg_cls_scores"]),
                                    torch.from_numpy(case["sg_iou_scores"]), torch.from_numpy(case["sg_mask_scores"]),
                                    case["num_ignored"])


def pack(inst, rle_decode, n_points):
    """list of pred dicts -> arrays (masks as sorted point lists in CSR form)."""
    label = np.array([int(d["label_id"]) for d in inst], np.int64)
    conf = np.array([d["conf"] for d in inst], np.float32)
    bbox = np.array([d["pred_bbox"] for d in inst], np.float32).reshape(len(inst), 6)
    pts, offs = [], [0]
    for d in inst:
        mask = rle_decode(d["pred_mask"]).astype(bool)
        assert mask.shape[0] == n_points
        pts.append(np.nonzero(mask)[0])
        offs.append(offs[-1] + pts[-1].size)
    pts = np.concatenate(pts) if pts else np.zeros(0, np.int64)
    return {"label_id": label, "conf": conf, "bbox": bbox, "mask_points": pts.astype(np.int32),
            "mask_offsets": np.asarray(offs, np.int32)}


def main():
    pg = reference_methods("minsu3d/model/pointgroup.py", "PointGroup", ("_get_nms_instances", "_get_pred_instances"))
    hs = reference_methods("minsu3d/model/hais.py", "HAIS", ("_get_pred_instances",))
    sg = reference_methods("minsu3d/model/softgroup.py", "SoftGroup", ("_get_pred_instances",))
    out = {}
    for ci, (n_points, n_obj, per_obj, seed) in enumerate(((4000, 6, 3, 1), (20000, 25, 4, 2), (500, 1, 1, 3))):
        rng = np.random.default_rng(seed)
        pidx, n_prop = make_proposals(rng, n_points, n_obj, per_obj)
        case = {"xyz": rng.uniform(-3, 3, (n_points, 3)).astype(np.float32),
                "scores": rng.normal(0.0, 2.0, (n_prop, 1)).astype(np.float32),
                "proposals_idx": pidx, "n_proposals": n_prop,
                "semantic_labels": rng.integers(0, 20, n_points).astype(np.int8),
                "mask_scores": rng.normal(0.3, 1.0, (pidx.shape[0], 1)).astype(np.float32),
                "num_ignored": 2, "nms_thr": 0.3, "score_thr": 0.09, "npoint_thr": 100 if n_points > 1000 else 10,
                "mask_thr": -0.5, "instance_classes": 18, "cls_thr": 0.05}
        # SoftGroup heads: class scores [P, 19] (softmax incl. background), IoU scores [P, 19], point mask scores [S, 19]
        case["sg_cls_scores"] = rng.normal(0, 2.0, (n_prop, 19)).astype(np.float32)
        case["sg_iou_scores"] = rng.uniform(-0.3, 1.3, (n_prop, 19)).astype(np.float32)
        case["sg_seed"] = seed  # sg_mask_scores = helpers.sg_mask_scores(seed, S): regenerated by the tests, not stored
        for k, v in case.items():
            out["c%d_in_%s" % (ci, k)] = np.asarray(v)
        case["sg_mask_scores"] = sg_mask_scores(seed, pidx.shape[0])
        # the reference takes scores and arg-maxes them; one-hot scores keep the fixture small
        case["semantic_scores"] = np.eye(20, dtype=np.float32)[case["semantic_labels"].astype(np.int64)]
        for name, res in (("pg", pack(run_pointgroup(case, pg), pg["rle_decode"], n_points)),
                          ("hais", pack(run_hais(case, hs), hs["rle_decode"], n_points)),
                          ("sg", pack(run_softgroup(case, sg), sg["rle_decode"], n_points))):
            for k, v in res.items():
                out["c%d_%s_%s" % (ci, name, k)] = v
            print("case", ci, name, "proposals", n_prop, "instances", res["label_id"].size)
    np.savez_compressed(os.path.join(HERE, "postproc_ref.npz"), **out)


if __name__ == "__main__":
    main()
