"""Seeded synthetic ScanNet-shaped scenes and the collated `data_dict` the hot path consumes.

Shapes/dtypes follow the reference's loader contract (SURVEY.md appendix B;
minsu3d/data/dataset/general_dataset.py:80-165, minsu3d/data/data_module.py:42-98): a scene is
N points sampled on the surfaces of a room (floor, two walls) and 6-12 axis-aligned boxes on a
jittered ~1.9 cm grid (about one point per 2 cm voxel, like ScanNet mesh vertices) with N(0, 4 mm)
noise along the normal; rgb ~ U(-1, 1); floor = class 0, wall = class 1 (the ignore classes),
boxes = classes 2..19 with one instance id per box.
"""
import numpy as np
import torch

VOXEL_SIZE = 0.02
IGNORE_CLASSES = (1, 2)  # config/data/scannetv2.yaml:18 (1-based ids of floor, wall)
NUM_CLASSES = 20
# config/data/scannetv2.yaml:25-33
POINT_NUM_AVG = [-1, -1, 3917, 12056, 2303, 8331, 3948, 3166, 5629, 11719,
                 1003, 3317, 4912, 10221, 3889, 4136, 2120, 945, 3967, 2589]
RADIUS_AVG = [-1., -1., 0.7047687683952325, 1.1732690381942337, 0.39644035821116036,
              1.011516629020215, 0.7260155292902369, 0.8674973999335017, 0.8374931435447094, 1.0454153869133096,
              0.32879464797430913, 1.1954566226966346, 0.8628817944400078, 1.0416287916782507, 0.6602697958671507,
              0.8541363897836871, 0.38055290598206537, 0.3011878752684007, 0.7420871812436316, 0.4474268644407741]


def _rect(rng, origin, u, v, spacing, noise):
    """Jittered grid on the rectangle origin + a*u + b*v, a,b in [0,1]."""
    lu, lv = np.linalg.norm(u), np.linalg.norm(v)
    nu, nv = max(int(lu / spacing), 1), max(int(lv / spacing), 1)
    a, b = np.meshgrid((np.arange(nu) + 0.5) / nu, (np.arange(nv) + 0.5) / nv, indexing="ij")
    a = a.reshape(-1) + rng.uniform(-0.3, 0.3, a.size) / nu
    b = b.reshape(-1) + rng.uniform(-0.3, 0.3, b.size) / nv
    normal = np.cross(u, v)
    normal = normal / np.linalg.norm(normal)
    pts = origin[None] + a[:, None] * u[None] + b[:, None] * v[None]
    return pts + rng.normal(0.0, noise, (pts.shape[0], 1)) * normal[None]


def make_scene(seed, n_points=100_000, area_scale=1.0):
    """One scene as numpy arrays: xyz f32 [N,3], rgb f32 [N,3], sem i16 [N], inst i16 [N]."""
    rng = np.random.default_rng(1000 + seed)
    s = float(np.sqrt(area_scale * n_points / 100_000.0))  # scale lengths so density stays ~constant
    W, D, H = 4.0 * s, 3.0 * s, 1.5 * s
    surfaces = [  # (origin, u, v, sem, inst)
        (np.array([0, 0, 0.0]), np.array([W, 0, 0.0]), np.array([0, D, 0.0]), 0, -1),
        (np.array([0, 0, 0.0]), np.array([W, 0, 0.0]), np.array([0, 0, H]), 1, -1),
        (np.array([0, 0, 0.0]), np.array([0, D, 0.0]), np.array([0, 0, H]), 1, -1),
    ]
    n_box = int(rng.integers(6, 13))
    for bi in range(n_box):
        sx, sy, sz = rng.uniform(0.35, 0.7, 3) * s
        x0, y0 = rng.uniform(0.2 * s, W - sx - 0.1 * s), rng.uniform(0.2 * s, D - sy - 0.1 * s)
        sem = 2 + int(rng.integers(0, 18))
        o = np.array([x0, y0, 0.0])
        ex, ey, ez = np.array([sx, 0, 0.0]), np.array([0, sy, 0.0]), np.array([0, 0, sz])
        for (oo, u, v) in ((o + ez, ex, ey), (o, ex, ez), (o + ey, ex, ez), (o, ey, ez), (o + ex, ey, ez)):
            surfaces.append((oo, u, v, sem, bi))
    area = sum(np.linalg.norm(np.cross(u, v)) for (_, u, v, _, _) in surfaces)
    spacing = np.sqrt(area / (n_points * 1.02))
    xyz, sem, inst = [], [], []
    for (o, u, v, sm, ins) in surfaces:
        p = _rect(rng, o, u, v, spacing, 0.004)
        xyz.append(p)
        sem.append(np.full(p.shape[0], sm, np.int16))
        inst.append(np.full(p.shape[0], ins, np.int16))
    xyz, sem, inst = np.concatenate(xyz), np.concatenate(sem), np.concatenate(inst)
    n = xyz.shape[0]
    if n >= n_points:
        keep = np.sort(rng.choice(n, n_points, replace=False))
    else:
        keep = np.sort(np.concatenate((np.arange(n), rng.choice(n, n_points - n, replace=True))))
    xyz, sem, inst = xyz[keep], sem[keep], inst[keep]
    xyz = xyz + rng.normal(0, 0.0005, xyz.shape)  # de-duplicate padded points
    xyz = (xyz - xyz.mean(0)).astype(np.float32)    # general_dataset.py:24 (scene-centred)
    rgb = rng.uniform(-1, 1, xyz.shape).astype(np.float32)
    # compact instance ids to 0..I-1 in order of appearance
    ids = np.unique(inst[inst >= 0])
    remap = -np.ones(int(inst.max()) + 2, np.int16)
    remap[ids] = np.arange(ids.size, dtype=np.int16)
    inst = np.where(inst >= 0, remap[np.clip(inst, 0, None)], -1).astype(np.int16)
    return {"xyz": xyz, "rgb": rgb, "sem_labels": sem, "instance_ids": inst}


def _inst_info(xyz, instance_ids, sem_labels):
    """general_dataset.py:56-78: per-point instance centre, per-instance size and class."""
    n_inst = int(instance_ids.max()) + 1 if instance_ids.size and instance_ids.max() >= 0 else 0
    center = np.zeros((xyz.shape[0], 3), np.float32)
    num_point, cls = [], []
    for i in range(n_inst):
        m = instance_ids == i
        center[m] = xyz[m].mean(0)
        num_point.append(int(m.sum()))
        c = int(sem_labels[m][0])
        cls.append(c - len(IGNORE_CLASSES) if (c + 1) not in IGNORE_CLASSES and c >= 0 else -1)
    return n_inst, center, np.asarray(num_point, np.int32), cls


def collate(scenes, device, quantize_device=None):
    """Batch scenes into the reference's data_dict (data_module.py:42-98) on `device`.

    Voxelisation = ME.utils.sparse_quantize(return_index, return_inverse, quantization_size=0.02)
    (general_dataset.py:159-163); it runs on the GPU when device is CUDA.
    """
    from ..MinkowskiEngine import utils as me_utils
    device = torch.device(device)
    qdev = quantize_device or ("cuda" if device.type == "cuda" else "cpu")
    out = {k: [] for k in ("point_xyz", "vert_batch_ids", "sem_labels", "instance_ids", "instance_center_xyz",
                           "instance_num_point", "voxel_xyz", "voxel_features", "voxel_point_map")}
    instance_offsets, instance_cls, total_inst, num_voxel = [0], [], 0, 0
    for b, sc in enumerate(scenes):
        xyz, rgb = sc["xyz"], sc["rgb"]
        inst = sc["instance_ids"].copy()
        n_inst, center, num_point, cls = _inst_info(xyz, inst, sc["sem_labels"])
        inst[inst != -1] += total_inst
        total_inst += n_inst
        instance_offsets.append(total_inst)
        instance_cls.extend(cls)
        feats = np.concatenate((rgb, xyz), axis=1)  # use_color + xyz (general_dataset.py:143-149)
        elastic = xyz - xyz.min(0)
        if qdev == "cuda":
            vx, vf, _, vmap = me_utils.sparse_quantize(torch.from_numpy(elastic).to(device),
                                                       torch.from_numpy(feats).to(device), return_index=True,
                                                       return_inverse=True, quantization_size=VOXEL_SIZE,
                                                       device="cuda")
        else:
            vx, vf, _, vmap = me_utils.sparse_quantize(elastic, feats, return_index=True, return_inverse=True,
                                                       quantization_size=VOXEL_SIZE)
            vx, vf = torch.from_numpy(vx), torch.from_numpy(vf)
        out["voxel_xyz"].append(vx)
        out["voxel_features"].append(vf)
        out["voxel_point_map"].append(vmap + num_voxel)
        num_voxel += vx.shape[0]
        out["point_xyz"].append(torch.from_numpy(xyz))
        out["vert_batch_ids"].append(torch.full((xyz.shape[0],), b, dtype=torch.uint8))
        out["sem_labels"].append(torch.from_numpy(sc["sem_labels"]))
        out["instance_ids"].append(torch.from_numpy(inst))
        out["instance_center_xyz"].append(torch.from_numpy(center))
        out["instance_num_point"].append(torch.from_numpy(num_point))
    data = {"scan_ids": ["synthetic_%04d" % i for i in range(len(scenes))]}
    for k in ("point_xyz", "vert_batch_ids", "sem_labels", "instance_ids", "instance_center_xyz",
              "instance_num_point", "voxel_point_map"):
        data[k] = torch.cat(out[k], dim=0).to(device)
    data["instance_offsets"] = torch.tensor(instance_offsets, dtype=torch.int32, device=device)
    data["instance_semantic_cls"] = torch.tensor(instance_cls, dtype=torch.int16, device=device)
    bcoords, bfeats = me_utils.sparse_collate(out["voxel_xyz"], out["voxel_features"])
    data["voxel_xyz"] = bcoords.to(device)
    data["voxel_features"] = bfeats.to(device)
    if device.type == "cuda":
        data["voxel_level_sizes"] = level_sizes(data["voxel_xyz"])
    return data


def level_sizes(voxel_xyz, levels=6):
    """{tensor stride: rows} of the strided coordinate maps of a voxelised batch -- a by-product of voxelisation that
    the loader hands to the model as host metadata (like the reference's loader hands over `voxel_point_map`), so
    that the MinkUNet forward needs no host read of device-side counts.  The model still builds every map on the
    device and validates these numbers there (ops.run_deferred_checks)."""
    from .. import ops
    return {int(s): int(oc.size(0)) for s, _, oc in ops.coord_pyramid(voxel_xyz.contiguous(), 1, levels)}


def make_batch(seeds, device, n_points=100_000):
    return collate([make_scene(s, n_points) for s in seeds], device)
