"""The reference loader's train-split sample pipeline on the GPU.

Reference: minsu3d/data/dataset/general_dataset.py:80-165 (`__getitem__`, train split) with
minsu3d/util/transform.py:65-98 (`elastic`, `crop`): augmentation matrix, colour jitter, two elastic distortions, shift
to the positive octant, crop, instance relabelling and statistics, voxelisation.  The reference runs this in four
DataLoader worker processes (scipy interpolation + CPU sparse_quantize); at > 150 scenes/s per GPU they cannot keep up.

Here the per-point work runs in libb2s kernels (csrc/augment.cu, the coordinate hash) on tensors that stay on the
device; the RANDOM DRAWS are inputs (`draw_augmentation()` makes them with numpy in the reference's call order), so the
result can be compared with the numpy restatement oracle/dataset_ref.py on shared draws (tests/test_gpu_loader.py).
Elastic distortion and voxel coordinates are computed in double precision like the reference's numpy / scipy code.
"""
import ctypes

import numpy as np
import torch

from .. import ops
from .._cabi import check, lib, ptr, stream
from ..MinkowskiEngine import utils as me_utils


def draw_augmentation(xyz_abs_max_after_affine=None, scene_xyz=None, voxel_size=0.02, rng=np.random):
    """The reference's random draws in its call order (general_dataset.py:28-41,95-98; transform.py:74).  The elastic
    noise volumes depend on the extent of the augmented scene, so the affine part is drawn first and the caller passes
    the scene's coordinates (numpy, float32)."""
    m = np.eye(3)
    m = np.matmul(m, np.eye(3) + rng.randn(3, 3) * 0.1)
    flip_m = np.eye(3)
    flip_m[0][0] *= rng.randint(0, 2) * 2 - 1
    m *= flip_m
    t = rng.rand() * 2 * np.pi
    c, s = np.cos(t), np.sin(t)
    m = np.matmul(m, np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]))
    draws = {"aug_matrix": m.astype(np.float32), "rgb_jitter": rng.randn(3) * 0.1, "elastic": [], "crop": []}
    return draws


def _as_c(arr, ctype):
    a = np.ascontiguousarray(arr)
    return a, a.ctypes.data_as(ctypes.POINTER(ctype))


def elastic_gpu(x, noise, gran, mag):
    """x [n,3] float64 CUDA tensor, modified in place; noise: numpy / tensor [3,b0,b1,b2] float32 (transform.py:65-85)."""
    nz = torch.as_tensor(noise, dtype=torch.float32).to(x.device).contiguous()
    tmp = torch.empty_like(nz)
    _, b0, b1, b2 = nz.shape
    check(lib().b2s_elastic_blur(ptr(nz), ptr(tmp), b0, b1, b2, stream()), "elastic_blur")
    check(lib().b2s_elastic_apply(ptr(x), ptr(nz), x.size(0), b0, b1, b2, float(gran), float(mag), stream()), "elastic_apply")
    return x


def train_sample_gpu(scene, draws, voxel_size=0.02, max_num_point=250000, full_scale=(128, 512), n_ignore=2,
                     elastic_noise=None):
    """scene: dict of CUDA tensors xyz f32 [N,3] (mean-centred), rgb f32 [N,3], sem_labels i16, instance_ids i16.
    draws: aug_matrix f32 [3,3], rgb_jitter [3], elastic = [noise0, noise1] (float32 [3,b0,b1,b2] each; when empty they
    are drawn here with numpy in the reference's order), crop = list of rand(3) draws consumed in order.
    Returns the per-sample dict of general_dataset.py:142-163 with CUDA tensors."""
    dev = scene["xyz"].device
    xyz, rgb = scene["xyz"].contiguous(), scene["rgb"].contiguous()
    n = xyz.size(0)
    m9, m9p = _as_c(np.asarray(draws["aug_matrix"], np.float32).reshape(9), ctypes.c_float)
    j3, j3p = _as_c(np.asarray(draws["rgb_jitter"], np.float64).astype(np.float32), ctypes.c_float)
    point_xyz = torch.empty_like(xyz)
    colors = torch.empty_like(rgb)
    check(lib().b2s_aug_affine(ptr(xyz), ptr(rgb), n, ctypes.cast(m9p, ctypes.c_void_p), ctypes.cast(j3p, ctypes.c_void_p),
                               ptr(point_xyz), ptr(colors), stream()), "aug_affine")
    # colours: the reference adds float64 jitter to a float32 array in place (`colors += randn(3) * 0.1`): the sum is
    # formed in double and rounded to float32 -- redo it exactly (the kernel's float32 add is the fast path)
    colors = (rgb.double() + torch.as_tensor(np.asarray(draws["rgb_jitter"], np.float64), device=dev)).float()
    scale = 1 / voxel_size
    x = (point_xyz * np.float32(scale)).double()  # `point_xyz * scale` is float32 in the reference
    plan = ((6 * scale // 50, 40 * scale / 50), (20 * scale // 50, 160 * scale / 50))
    for i, (gran, mag) in enumerate(plan):
        if i < len(draws["elastic"]):
            noise = draws["elastic"][i]
        else:  # draw like transform.py:73-74 (needs the extent: one host read)
            bb = (np.abs(x.float().cpu().numpy() if i == 0 else x.cpu().numpy()).max(0) // gran + 3).astype(np.int32)
            noise = np.stack([np.random.randn(bb[0], bb[1], bb[2]).astype(np.float32) for _ in range(3)])
            draws["elastic"].append(noise)
        elastic_gpu(x, noise, gran, mag)
    x -= x.min(dim=0).values
    # ---- crop (general_dataset.py:110-134, transform.py:88-98) ---------------------------------------------------
    sem, inst = scene["sem_labels"], scene["instance_ids"]
    valid = None
    if n > max_num_point:
        crop_draws = iter(draws["crop"])
        out = torch.empty_like(x)
        vmask = torch.empty(n, dtype=torch.uint8, device=dev)
        d_count = torch.zeros(1, dtype=torch.int32, device=dev)
        pc_range = (x.max(dim=0).values - x.min(dim=0).values).cpu().numpy()
        tries, count, accepted = 20, 0, False
        while tries > 0:
            rng_max = np.full(3, full_scale[1], dtype=np.uint16)
            # first test of transform.crop: no offset, only `min >= 0` (range = +inf)
            off, offp = _as_c(np.zeros(3), ctypes.c_double)
            big, bigp = _as_c(np.full(3, np.inf), ctypes.c_double)
            check(lib().b2s_crop_test(ptr(x), n, ctypes.cast(offp, ctypes.c_void_p), ctypes.cast(bigp, ctypes.c_void_p),
                                      ptr(out), ptr(vmask), ptr(d_count), stream()), "crop_test")
            count = int(d_count.item())
            while count > max_num_point:
                r = np.asarray(next(crop_draws), np.float64)
                off, offp = _as_c(np.clip(rng_max - pc_range + 0.001, None, 0) * r, ctypes.c_double)
                rg, rgp = _as_c(rng_max.astype(np.float64), ctypes.c_double)
                check(lib().b2s_crop_test(ptr(x), n, ctypes.cast(offp, ctypes.c_void_p), ctypes.cast(rgp, ctypes.c_void_p),
                                          ptr(out), ptr(vmask), ptr(d_count), stream()), "crop_test")
                count = int(d_count.item())
                rng_max[:2] -= 32
            valid = vmask.bool()
            if count >= max_num_point // 2 and bool((sem[valid] != -1).any()) and bool((inst[valid] != -1).any()):
                x = out.clone()
                accepted = True
                break
            tries -= 1
        if not accepted:
            raise RuntimeError("Over-cropped!")
    if valid is not None:
        keep = torch.nonzero(valid).view(-1)
        x, point_xyz, colors, sem = x[keep], point_xyz[keep], colors[keep], sem[keep]
        inst = inst[keep].clone()
        # _get_cropped_inst_ids (:43-53): the highest id moves into every gap; the mapping is computed on the host
        # from the (small) presence histogram and applied on the device
        present = torch.bincount((inst[inst >= 0]).long()).cpu().numpy() > 0 if bool((inst >= 0).any()) else np.zeros(0, bool)
        ids = {i for i in range(present.size) if present[i]}
        mapping = {}
        j = 0
        while ids and j < max(ids):
            if j not in ids:
                top = max(ids)
                ids.remove(top)
                ids.add(j)
                mapping[top] = j
            j += 1
        if mapping:
            table = torch.arange(present.size, dtype=torch.int16, device=dev)
            # chains (an id moved twice) resolve by following the mapping to its end
            for src in list(mapping):
                dst = mapping[src]
                while dst in mapping:
                    dst = mapping[dst]
                mapping[src] = dst
            for orig in range(present.size):
                cur = orig
                while cur in mapping:
                    cur = mapping[cur]
                table[orig] = cur
            inst = torch.where(inst >= 0, table[inst.clamp(min=0).long()], inst)
    x = x / scale
    # ---- instance statistics (general_dataset.py:55-78) ----------------------------------------------------------
    n_pts = point_xyz.size(0)
    fg = inst >= 0
    n_inst = int(inst.max().item()) + 1 if bool(fg.any()) else 0
    center = torch.zeros((n_pts, 3), dtype=torch.float32, device=dev)
    num_point = torch.zeros(n_inst, dtype=torch.int32, device=dev)
    cls = torch.full((n_inst,), -1, dtype=torch.int16, device=dev)
    if n_inst:
        ids = inst[fg].long()
        cnt = torch.bincount(ids, minlength=n_inst)
        sums = torch.zeros((n_inst, 3), dtype=torch.float64, device=dev).index_add_(0, ids, point_xyz[fg].double())
        mean = (sums / cnt.clamp(min=1)[:, None]).float()
        center[fg] = mean[ids]
        present = cnt > 0  # np.unique: only ids that occur; relabelling made them dense
        num_point = cnt[present].int()
        first = torch.full((n_inst,), n_pts, dtype=torch.int64, device=dev).scatter_reduce_(
            0, ids, torch.nonzero(fg).view(-1), reduce="amin")
        c = sem[first[present].clamp(max=n_pts - 1)]
        cls = torch.where(c != -1, c - n_ignore, c).to(torch.int16)
        n_inst = int(present.sum().item())
    feats = torch.cat((colors, point_xyz), dim=1)
    vx, vf, _, vmap = me_utils.sparse_quantize(x, feats, return_index=True, return_inverse=True,
                                               quantization_size=voxel_size, device="cuda")
    return {"point_xyz": point_xyz, "sem_labels": sem, "instance_ids": inst, "num_instance": n_inst,
            "instance_center_xyz": center, "instance_num_point": num_point, "instance_semantic_cls": cls,
            "voxel_xyz": vx[:, 1:].contiguous() if vx.size(1) == 4 else vx, "voxel_features": vf, "voxel_point_map": vmap,
            "point_xyz_elastic": x}
