"""numpy restatement of the reference's train-split sample pipeline -- TEST INFRASTRUCTURE ONLY.

Follows minsu3d/data/dataset/general_dataset.py:28-165 (`_get_augmentation_matrix`, `__getitem__`) and
minsu3d/util/transform.py:6-25,65-98 (`jitter`, `flip`, `rotz`, `elastic`, `crop`) statement by statement and calls
`np.random` in the SAME ORDER, so `np.random.seed(s)` before either this function or the reference's own code gives
identical draws (tests/test_cpu_oracle_and_host.py pins it against the reference's functions where they are available).
Every draw is also recorded in `draws`, which is what the GPU path (minsu3d_b200/harness/gpu_loader.py) consumes."""
import numpy as np
import scipy.interpolate
import scipy.ndimage


def elastic(x, gran, mag, draws):
    """transform.py:65-85."""
    blur0 = np.ones((3, 1, 1), dtype=np.float32) / 3
    blur1 = np.ones((1, 3, 1), dtype=np.float32) / 3
    blur2 = np.ones((1, 1, 3), dtype=np.float32) / 3
    bb = (np.abs(x).max(0) // gran + 3).astype(np.int32)
    noise = [np.random.randn(bb[0], bb[1], bb[2]).astype(np.float32) for _ in range(3)]
    draws.append(np.stack(noise))
    for blur in (blur0, blur1, blur2, blur0, blur1, blur2):
        noise = [scipy.ndimage.convolve(n, blur, mode="constant", cval=0) for n in noise]
    ax = [np.linspace(-(b - 1) * gran, (b - 1) * gran, b) for b in bb]
    interp = [scipy.interpolate.RegularGridInterpolator(ax, n, bounds_error=0, fill_value=0) for n in noise]
    return x + np.hstack([i(x)[:, None] for i in interp]) * mag


def crop(pc, max_num_point, scale, draws):
    """transform.py:88-98."""
    pc_offset = pc.copy()
    valid_idxs = pc_offset.min(1) >= 0
    max_pc_range = np.full(shape=3, fill_value=scale, dtype=np.uint16)
    pc_range = pc.max(0) - pc.min(0)
    while np.count_nonzero(valid_idxs) > max_num_point:
        r = np.random.rand(3)
        draws.append(r)
        offset = np.clip(max_pc_range - pc_range + 0.001, None, 0) * r
        pc_offset = pc + offset
        valid_idxs = np.logical_and(pc_offset.min(1) >= 0, np.all(pc_offset < max_pc_range, axis=1))
        max_pc_range[:2] -= 32
    return pc_offset, valid_idxs


def cropped_inst_ids(instance_ids, valid_idxs):
    """general_dataset.py:43-53."""
    instance_ids = instance_ids[valid_idxs]
    j = 0
    while j < instance_ids.max():
        if np.count_nonzero(instance_ids == j) == 0:
            instance_ids[instance_ids == instance_ids.max()] = j
        j += 1
    return instance_ids


def inst_info(xyz, instance_ids, sem_labels, n_ignore):
    """general_dataset.py:55-78."""
    unique_ids = np.unique(instance_ids)
    unique_ids = unique_ids[unique_ids != -1]
    center = np.empty((xyz.shape[0], 3), np.float32)
    num_point = []
    cls = np.full(unique_ids.shape[0], -1, np.int16)
    for index, i in enumerate(unique_ids):
        idx = np.where(instance_ids == i)[0]
        center[idx] = xyz[idx].mean(0)
        num_point.append(idx.size)
        c = sem_labels[idx[0]]
        cls[index] = c - n_ignore if c != -1 else c
    return unique_ids.shape[0], center, np.array(num_point, np.int32), cls


def train_sample(scene, voxel_size=0.02, max_num_point=250000, full_scale=(128, 512), n_ignore=2, use_color=True):
    """general_dataset.py:80-165 for split == "train" (all augmentations on, config/data/base.yaml:11-16).
    scene: dict(xyz f32 [N,3] (already mean-centred, :24), rgb f32 in [-1,1] (:25), sem_labels, instance_ids).
    Returns (data dict, draws dict)."""
    draws = {"elastic": [], "crop": []}
    point_xyz = scene["xyz"].astype(np.float32)
    colors = scene["rgb"].astype(np.float32).copy()
    instance_ids = scene["instance_ids"].astype(np.int16)
    sem_labels = scene["sem_labels"].astype(np.int16)
    # _get_augmentation_matrix (:28-41)
    m = np.eye(3)
    jit = np.random.randn(3, 3)
    m = np.matmul(m, np.eye(3) + jit * 0.1)
    f = np.random.randint(0, 2)
    flip_m = np.eye(3)
    flip_m[0][0] *= f * 2 - 1
    m *= flip_m
    t = np.random.rand() * 2 * np.pi
    c, s = np.cos(t), np.sin(t)
    m = np.matmul(m, np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]))
    aug = m.astype(np.float32)
    point_xyz = np.matmul(point_xyz, aug)
    rgb_jit = np.random.randn(3) * 0.1
    colors += rgb_jit
    draws.update(aug_matrix=aug, rgb_jitter=rgb_jit)
    # elastic (:100-107)
    scale = 1 / voxel_size
    e = elastic(point_xyz * scale, 6 * scale // 50, 40 * scale / 50, draws["elastic"])
    e = elastic(e, 20 * scale // 50, 160 * scale / 50, draws["elastic"])
    e -= e.min(axis=0)
    # crop (:110-134)
    valid_idxs = np.ones(point_xyz.shape[0], dtype=bool)
    if valid_idxs.shape[0] > max_num_point:
        max_tries, count = 20, 0
        while max_tries > 0:
            tmp, valid_idxs = crop(e, max_num_point, full_scale[1], draws["crop"])
            count = np.count_nonzero(valid_idxs)
            if count >= (max_num_point // 2) and np.any(sem_labels[valid_idxs] != -1) and np.any(instance_ids[valid_idxs] != -1):
                e = tmp
                break
            max_tries -= 1
        if count < (max_num_point // 2) or np.all(sem_labels[valid_idxs] == -1) and np.all(instance_ids[valid_idxs] == -1):
            raise Exception("Over-cropped!")
    e = e[valid_idxs]
    point_xyz = point_xyz[valid_idxs]
    colors = colors[valid_idxs]
    sem_labels = sem_labels[valid_idxs]
    instance_ids = cropped_inst_ids(instance_ids, valid_idxs)
    e /= (1 / voxel_size)
    n_inst, center, num_point, cls = inst_info(point_xyz, instance_ids, sem_labels, n_ignore)
    feats = np.concatenate((colors, point_xyz), axis=1) if use_color else point_xyz
    # ME.utils.sparse_quantize(coordinates=e, features, return_index, return_inverse, quantization_size) (:159-163)
    dc = np.floor(e / voxel_size).astype(np.int32)
    _, first, inv = np.unique(dc, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    uniq = first[order]
    data = {"point_xyz": point_xyz, "sem_labels": sem_labels, "instance_ids": instance_ids,
            "num_instance": np.array(n_inst, np.int32), "instance_center_xyz": center, "instance_num_point": num_point,
            "instance_semantic_cls": cls, "voxel_xyz": dc[uniq], "voxel_features": feats[uniq].astype(np.float32),
            "voxel_point_map": rank[inv.reshape(-1)].astype(np.int64), "point_xyz_elastic": e, "valid_idxs": valid_idxs}
    return data, draws
