"""Generates tests/golden/*.npz from the REFERENCE's own compiled COMMON_OPS (oracle/_ref).

Run where oracle/_ref exists:
    python tests/golden/make_golden.py            # CPU entry points (pg/sg BFS, hierarchical_aggregation w/o set aggr)
    python tests/golden/make_golden.py --gpu      # + the reference CUDA kernels (needs a GPU: run under gpurun,
                                                  #   then copy gpurun_out/golden_gpu.npz to tests/golden/)
Inputs are seeded (tests/helpers.py) and stored next to the outputs, so the fixtures are self-contained.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import clustered_points  # noqa: E402
from oracle import build_ref  # noqa: E402

POINT_NUM_AVG = [-1, -1, 3917, 12056, 2303, 8331, 3948, 3166, 5629, 11719, 1003, 3317, 4912, 10221, 3889, 4136,
                 2120, 945, 3967, 2589]
RADIUS_AVG = [-1., -1., 0.7047687683952325, 1.1732690381942337, 0.39644035821116036, 1.011516629020215,
              0.7260155292902369, 0.8674973999335017, 0.8374931435447094, 1.0454153869133096, 0.32879464797430913,
              1.1954566226966346, 0.8628817944400078, 1.0416287916782507, 0.6602697958671507, 0.8541363897836871,
              0.38055290598206537, 0.3011878752684007, 0.7420871812436316, 0.4474268644407741]


def E(dtype=torch.int32):
    return torch.empty(0, dtype=dtype)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def brute_ballquery(xyz, bidx, offs, radius, cap=1000):
    """numpy restatement used ONLY to build the CSR input of the CPU goldens (float32 arithmetic)."""
    return oracle.ballquery(xyz, bidx, offs, radius)


def cpu_goldens(ref):
    out = {}
    rng = np.random.default_rng(2024)
    # KAT from SURVEY.md section 8(c), computed by the reference binary
    nbrs = [[0, 1], [0, 1, 2], [1, 2], [3, 4], [3, 4], [5]]
    idx = np.asarray(sum(nbrs, []), np.int32)
    sl = np.asarray([[sum(len(x) for x in nbrs[:i]), len(nbrs[i])] for i in range(6)], np.int32)
    lab = np.asarray([3, 3, 3, 4, 4, 4], np.int16)
    ci, co = E(), E()
    ref.pg_bfs_cluster(t(lab), t(idx), t(sl), ci, co, 6, 2)
    out.update(kat_idx=idx, kat_sl=sl, kat_lab=lab, kat_ci=ci.numpy().copy(), kat_co=co.numpy().copy())
    for tag, n, n_obj, collapse in (("sparse", 6000, 6, False), ("dense", 9000, 2, True)):
        xyz, lab, bidx, offs = clustered_points(rng, n, n_obj=n_obj, collapse=collapse)
        lab = lab.copy()
        lab[rng.integers(0, lab.size, lab.size // 25)] = 9
        idx, sl = brute_ballquery(xyz, bidx, offs, 0.03)
        ci, co = E(), E()
        ref.pg_bfs_cluster(t(lab), t(idx), t(sl), ci, co, sl.shape[0], 20)
        si, so = E(), E()
        ref.sg_bfs_cluster(torch.tensor(POINT_NUM_AVG, dtype=torch.float32), t(idx), t(sl), si, so, sl.shape[0], 0.05, 10)
        fr = [E(), E(), E(torch.float32)]
        k = [E(), E(), E(torch.float32)]
        p = [E(), E(), E(torch.float32)]
        pp = [E(), E()]
        ref.hierarchical_aggregation(t(lab), t(xyz), t(bidx), t(idx), t(sl), *fr, *k, *p, *pp,
                                     torch.tensor(POINT_NUM_AVG, dtype=torch.float32),
                                     torch.tensor(RADIUS_AVG, dtype=torch.float32), sl.shape[0], 0, -1)
        out.update({tag + "_xyz": xyz, tag + "_lab": lab, tag + "_bidx": bidx, tag + "_offs": offs,
                    tag + "_idx": idx, tag + "_sl": sl,
                    tag + "_pg_ci": ci.numpy().copy(), tag + "_pg_co": co.numpy().copy(),
                    tag + "_sg_ci": si.numpy().copy(), tag + "_sg_co": so.numpy().copy(),
                    tag + "_ha_kept_i": k[0].numpy().copy(), tag + "_ha_kept_o": k[1].numpy().copy(),
                    tag + "_ha_kept_c": k[2].numpy().copy(), tag + "_ha_prim_i": p[0].numpy().copy(),
                    tag + "_ha_prim_o": p[1].numpy().copy(), tag + "_ha_prim_c": p[2].numpy().copy()})
    np.savez_compressed(os.path.join(HERE, "common_ops_cpu.npz"), **out)
    print("wrote common_ops_cpu.npz", {k: v.shape for k, v in out.items() if k.endswith(("_ci", "_co"))})


def gpu_goldens(ref, path):
    out = {}
    rng = np.random.default_rng(4048)
    xyz, lab, bidx, offs = clustered_points(rng, 8000, n_obj=6)
    n = xyz.shape[0]
    d = lambda a: t(a).cuda()  # noqa: E731
    ma = 300
    idx = torch.zeros(n * ma, dtype=torch.int32, device="cuda")
    sl = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    na = ref.ballquery_batch_p(d(xyz), d(bidx), d(offs), idx, sl, n, ma, 0.03)
    idx, sl = idx[:na].cpu().numpy(), sl.cpu().numpy()
    order = np.argsort(sl[:, 0], kind="stable")  # canonical CSR: lists in point order
    lists = [idx[s:s + l] for s, l in sl]
    out.update(bq_xyz=xyz, bq_bidx=bidx, bq_offs=offs, bq_len=sl[:, 1].copy(), bq_idx=np.concatenate(lists))
    lens = rng.integers(1, 900, 60)
    so = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
    x = np.round(rng.standard_normal((so[-1], 16)), 2).astype(np.float32)
    for kind in ("mean", "min", "max"):
        o = torch.zeros((60, 16), device="cuda")
        getattr(ref, "sec_" + kind)(d(x), d(so), o, 60, 16)
        out["sec_" + kind] = o.cpu().numpy()
    o = torch.zeros((60, 16), device="cuda")
    a = torch.zeros((60, 16), dtype=torch.int32, device="cuda")
    ref.roipool_fp(d(x), d(so), o, a, 60, 16)
    g = torch.zeros((60, 16), device="cuda")
    ref.global_avg_pool_fp(d(x), d(so), g, 60, 16)
    out.update(seg_x=x, seg_offs=so, roipool_out=o.cpu().numpy(), roipool_arg=a.cpu().numpy(), gap=g.cpu().numpy())
    inst = rng.integers(-1, 23, 20_000).astype(np.int16)
    inst_num = np.bincount(inst[inst >= 0], minlength=23).astype(np.int32)
    plen = rng.integers(30, 800, 40)
    po = np.concatenate(([0], np.cumsum(plen))).astype(np.int32)
    pidx = np.concatenate([np.sort(rng.choice(20_000, l, replace=False)) for l in plen]).astype(np.int32)
    for p in range(40):
        sel = np.nonzero(inst == (p % 23))[0][:plen[p] // 2]
        pidx[po[p]:po[p] + sel.size] = sel
    cls = rng.integers(-1, 18, 23).astype(np.int16)
    sc = rng.uniform(0, 1, pidx.size).astype(np.float32)
    iou = torch.zeros((40, 23), device="cuda")
    ref.get_iou(d(pidx), d(po), d(inst), d(inst_num), iou, 23, 40)
    ioup = torch.zeros((40, 23), device="cuda")
    ref.get_mask_iou_on_pred(d(pidx), d(po), d(inst), d(inst_num), ioup, 23, 40, d(sc))
    ml = torch.zeros(pidx.size, dtype=torch.bool, device="cuda")
    mm = torch.zeros(pidx.size, dtype=torch.bool, device="cuda")
    ref.get_mask_label(d(pidx), d(po), d(inst), d(cls), iou, 23, 40, -1, 0.25, ml, mm)
    out.update(iou_pidx=pidx, iou_poff=po, iou_inst=inst, iou_inst_num=inst_num, iou_cls=cls, iou_scores=sc,
               iou=iou.cpu().numpy(), iou_pred=ioup.cpu().numpy(), mask_label=ml.cpu().numpy(),
               mask_label_mask=mm.cpu().numpy())
    np.savez_compressed(path, **out)
    print("wrote", path)


def gpu_golden_clusters_voxelize(ref, path):
    """The reference's torch expression sequence of clusters_voxelization (general_model.py:154-184) executed on the
    GPU with the reference's own sec_mean / sec_min / sec_max kernels (the function itself cannot be imported: its
    module imports MinkowskiEngine)."""
    rng = np.random.default_rng(777)
    n, n_cluster, scale, shape = 30_000, 48, 50, 14
    coords = (rng.random((n, 3)) * 8.0).astype(np.float32)
    sizes = rng.integers(1, 1500, n_cluster)
    sizes[0] = 1
    offs = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    pts = np.concatenate([rng.integers(0, n, 1) + rng.integers(0, 300, s) for s in sizes]) % n
    idx = np.stack((np.repeat(np.arange(n_cluster), sizes), pts), 1).astype(np.int64)
    rand = rng.random((2, 3)).astype(np.float32)
    d = lambda a: t(a).cuda()  # noqa: E731
    clusters_idx, clusters_offset, dcoords, drand = d(idx), d(offs), d(coords), d(rand)

    def sec(kind, x):
        o = torch.zeros((n_cluster, 3), device="cuda")
        getattr(ref, "sec_" + kind)(x.contiguous(), clusters_offset, o, n_cluster, 3)
        return o

    batch_idx = clusters_idx[:, 0]
    cc = dcoords[clusters_idx[:, 1]]
    mean = sec("mean", cc)
    cc = cc - torch.index_select(mean, 0, batch_idx)
    cmin, cmax = sec("min", cc), sec("max", cc)
    cscale = 1 / ((cmax - cmin) / shape).max(1)[0] - 0.01
    cscale = torch.clamp(cscale, min=None, max=scale)
    min_xyz, max_xyz = cmin * cscale[:, None], cmax * cscale[:, None]
    cc = cc * torch.index_select(cscale, 0, batch_idx)[:, None]
    rg = max_xyz - min_xyz
    offset = -min_xyz + torch.clamp(shape - rg - 0.001, min=0) * drand[0]
    offset += torch.clamp(shape - rg + 0.001, max=0) * drand[1]
    cc = cc + torch.index_select(offset, 0, batch_idx)
    cc = cc.int()
    batched = torch.cat((clusters_idx[:, 0].unsqueeze(-1).int(), cc), dim=1)
    np.savez_compressed(path, idx=idx, offs=offs, coords=coords, rand=rand, scale=np.float32(scale),
                        shape=np.int32(shape), batched_xyz=batched.cpu().numpy())
    print("wrote", path)


if __name__ == "__main__":
    ref = build_ref.load()
    assert ref is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    if "--gpu" in sys.argv:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        gpu_goldens(ref, os.path.join(ROOT, "gpurun_out", "golden_gpu.npz"))
        gpu_golden_clusters_voxelize(ref, os.path.join(ROOT, "gpurun_out", "golden_clusters_voxelize_gpu.npz"))
    else:
        cpu_goldens(ref)
