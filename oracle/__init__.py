"""oracle -- CPU restatement of the reference algorithms.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package; minsu3d_b200 (the product) never does.

* oracle.c / this module: plain-C restatement (numpy in, numpy out) of
    - the COMMON_OPS algorithms (minsu3d/common_ops/src/**), PINNED against the reference's own
      compiled extension (oracle/_ref, built from /root/reference by oracle/build_ref.py) and the
      golden vectors it generated (tests/golden/);
    - the MinkowskiEngine CPU-backend algorithm.  MinkowskiEngine is an un-vendored pip dependency
      (README.md:44-46,73; not in /root/reference, not installable offline): PARITY UNPINNED against
      the real binary; cross-checked against torch conv3d and the dictionary restatement me_ref.py.
* me_ref.py: pure-Python dictionary restatement of the ME semantics (small cases only).
* build_ref.py: recipe that compiles the unmodified reference COMMON_OPS into oracle/_ref/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB = os.path.join(_OUT_DIR, "liboracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -fopenmp -shared oracle.c -> oracle/_build/liboracle.so"""
    if (not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC)):
        return _LIB
    os.makedirs(_OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
                           _SRC, "-o", _LIB, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_ballquery.restype = ctypes.c_int64
        _lib.orc_ha_concat.restype = ctypes.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


# ---- Part A: COMMON_OPS --------------------------------------------------------------------
def ballquery(xyz, batch_idxs, batch_offsets, radius):
    """-> (idx [nActive] i32, start_len [n,2] i32), canonical CSR (bfs_cluster.cu:15-60)."""
    xyz, batch_idxs, batch_offsets = _c(xyz, np.float32), _c(batch_idxs, np.uint8), _c(batch_offsets, np.int32)
    n = xyz.shape[0]
    start_len = np.zeros((n, 2), np.int32)
    r = ctypes.c_float(radius)
    total = lib().orc_ballquery(_p(xyz), _p(batch_idxs), _p(batch_offsets), n, r, None, _p(start_len))
    idx = np.zeros(total, np.int32)
    lib().orc_ballquery(_p(xyz), _p(batch_idxs), _p(batch_offsets), n, r, _p(idx), _p(start_len))
    return idx, start_len


def bfs_cluster(labels, nbr_idx, start_len, mode, thr_i=0, thr_f=0.0, point_num_avg=None, group=0,
                coords=None, batch_idxs=None):
    """-> (cluster_idxs [S,2], cluster_offsets [nC+1], centers [nC,5] or None)."""
    labels = _c(labels, np.int16)
    nbr_idx, start_len = _c(nbr_idx, np.int32), _c(start_len, np.int32)
    n = start_len.shape[0]
    cluster_idxs = np.zeros((max(n, 1), 2), np.int32)
    cluster_offsets = np.zeros(n + 1, np.int32)
    coords, batch_idxs = _c(coords, np.float32), _c(batch_idxs, np.uint8)
    centers = np.zeros((max(n, 1), 5), np.float32) if coords is not None else None
    pna = _c(point_num_avg, np.float32)
    total = ctypes.c_int64(0)
    nc = lib().orc_bfs_cluster(_p(labels), _p(nbr_idx), _p(start_len), n, mode, int(thr_i), ctypes.c_float(thr_f),
                               _p(pna), group, _p(coords), _p(batch_idxs), _p(cluster_idxs), _p(cluster_offsets),
                               _p(centers), ctypes.byref(total))
    return (cluster_idxs[:total.value].copy(), cluster_offsets[:nc + 1].copy(),
            None if centers is None else centers[:nc].copy())


def pg_bfs_cluster(labels, nbr_idx, start_len, threshold):
    ci, co, _ = bfs_cluster(labels, nbr_idx, start_len, 0, thr_i=threshold)
    return ci, co


def sg_bfs_cluster(class_numpoint_mean, nbr_idx, start_len, threshold, class_id):
    mean = np.float32(class_numpoint_mean[class_id])
    thr = np.float32(threshold) if mean == -1 else np.float32(threshold) * mean
    ci, co, _ = bfs_cluster(None, nbr_idx, start_len, 1, thr_f=float(thr))
    return ci, co


def ha_assign(frag_centers, prim_centers, prim_offsets, radius_avg):
    fc, pc = _c(frag_centers, np.float32), _c(prim_centers, np.float32)
    po, ra = _c(prim_offsets, np.int32), _c(radius_avg, np.float32)
    assign = np.full(max(fc.shape[0], 1), -1, np.int32)
    lib().orc_ha_assign(_p(fc), fc.shape[0], _p(pc), _p(po), pc.shape[0], _p(ra), _p(assign))
    return assign[:fc.shape[0]]


def ha_concat(frag_idxs, frag_offsets, prim_idxs, prim_offsets, assign):
    fi, fo = _c(frag_idxs, np.int32), _c(frag_offsets, np.int32)
    pi, po, asg = _c(prim_idxs, np.int32), _c(prim_offsets, np.int32), _c(assign, np.int32)
    n_prim = po.shape[0] - 1
    out = np.zeros((fi.shape[0] + pi.shape[0] + 1, 2), np.int32)
    off = np.zeros(n_prim + 1, np.int32)
    tot = lib().orc_ha_concat(_p(fi), _p(fo), fo.shape[0] - 1, _p(pi), _p(po), n_prim, _p(asg), _p(out), _p(off))
    return out[:tot].copy(), off


def hierarchical_aggregation(labels, coord_shift, nbr_idx, start_len, batch_idxs, using_set_aggr,
                             point_num_avg, radius_avg):
    """Restates hais_ops.py:8-73 + hierarchical_aggregation.cpp:108-183 end to end."""
    kept_i, kept_o, _ = bfs_cluster(labels, nbr_idx, start_len, 2, point_num_avg=point_num_avg, group=1,
                                    coords=coord_shift, batch_idxs=batch_idxs)
    prim_i, prim_o, prim_c = bfs_cluster(labels, nbr_idx, start_len, 2, point_num_avg=point_num_avg, group=2,
                                         coords=coord_shift, batch_idxs=batch_idxs)
    if using_set_aggr and prim_o.shape[0] > 1:
        frag_i, frag_o, frag_c = bfs_cluster(labels, nbr_idx, start_len, 2, point_num_avg=point_num_avg, group=3,
                                             coords=coord_shift, batch_idxs=batch_idxs)
        assign = ha_assign(frag_c, prim_c, prim_o, radius_avg)
        prim_i, prim_o = ha_concat(frag_i, frag_o, prim_i, prim_o, assign)
    ci, co = kept_i, kept_o
    if prim_i.shape[0] != 0:
        prim_i = prim_i.copy()
        prim_i[:, 0] += co.shape[0] - 1
        ci = np.concatenate((ci, prim_i), 0)
        co = np.concatenate((co, prim_o[1:] + co[-1]))
    return ci, co


def sec(kind, inp, offsets):
    inp, offsets = _c(inp, np.float32), _c(offsets, np.int32)
    out = np.zeros((offsets.shape[0] - 1, inp.shape[1]), np.float32)
    lib().orc_sec(_p(inp), _p(offsets), _p(out), out.shape[0], inp.shape[1], {"mean": 0, "min": 1, "max": 2}[kind])
    return out


def roipool_fp(feats, offsets):
    feats, offsets = _c(feats, np.float32), _c(offsets, np.int32)
    out = np.zeros((offsets.shape[0] - 1, feats.shape[1]), np.float32)
    arg = np.zeros(out.shape, np.int32)
    lib().orc_roipool_fp(_p(feats), _p(offsets), _p(out), _p(arg), out.shape[0], feats.shape[1])
    return out, arg


def roipool_bp(n_rows, maxidx, d_out):
    maxidx, d_out = _c(maxidx, np.int32), _c(d_out, np.float32)
    d_feats = np.zeros((n_rows, d_out.shape[1]), np.float32)
    lib().orc_roipool_bp(_p(d_feats), _p(maxidx), _p(d_out), d_out.shape[0], d_out.shape[1])
    return d_feats


def gap_fp(feats, offsets):
    feats, offsets = _c(feats, np.float32), _c(offsets, np.int32)
    out = np.zeros((offsets.shape[0] - 1, feats.shape[1]), np.float32)
    lib().orc_gap_fp(_p(feats), _p(offsets), _p(out), out.shape[0], feats.shape[1])
    return out


def gap_bp(n_rows, offsets, d_out):
    offsets, d_out = _c(offsets, np.int32), _c(d_out, np.float32)
    d_feats = np.zeros((n_rows, d_out.shape[1]), np.float32)
    lib().orc_gap_bp(_p(d_feats), _p(offsets), _p(d_out), d_out.shape[0], d_out.shape[1])
    return d_feats


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, mask_scores=None):
    pi, po = _c(proposals_idx, np.int32), _c(proposals_offset, np.int32)
    il, ip = _c(instance_labels, np.int16), _c(instance_pointnum, np.int32)
    ms = _c(mask_scores, np.float32)
    iou = np.zeros((po.shape[0] - 1, ip.shape[0]), np.float32)
    lib().orc_get_iou(_p(pi), _p(po), _p(il), _p(ip), _p(ms), _p(iou), ip.shape[0], po.shape[0] - 1)
    return iou


def get_mask_label(proposals_idx, proposals_offset, instance_labels, instance_cls, iou, ignored_label, iou_thr):
    pi, po = _c(proposals_idx, np.int32), _c(proposals_offset, np.int32)
    il, ic, iou = _c(instance_labels, np.int16), _c(instance_cls, np.int16), _c(iou, np.float32)
    ml = np.zeros(pi.shape[0], np.uint8)
    mm = np.zeros(pi.shape[0], np.uint8)
    lib().orc_get_mask_label(_p(pi), _p(po), _p(il), _p(ic), _p(iou), iou.shape[1], po.shape[0] - 1,
                             int(ignored_label), ctypes.c_float(iou_thr), _p(ml), _p(mm))
    return ml.astype(bool), mm.astype(bool)


# ---- Part B: MinkowskiEngine CPU-backend restatement -----------------------------------------
def coord_unique(coords, quant=1):
    """-> (unique_idx [m], inverse [n], out_coords [m,4]) first-occurrence order."""
    coords = _c(coords, np.int32)
    n = coords.shape[0]
    ui = np.zeros(max(n, 1), np.int32)
    inv = np.zeros(max(n, 1), np.int32)
    oc = np.zeros((max(n, 1), 4), np.int32)
    m = lib().orc_coord_unique(_p(coords), n, int(quant), _p(ui), _p(inv), _p(oc))
    return ui[:m].copy(), inv[:n].copy(), oc[:m].copy()


def kernel_map(in_coords, out_coords, ksize, dil):
    ic, oc = _c(in_coords, np.int32), _c(out_coords, np.int32)
    nbr = np.zeros((oc.shape[0], ksize ** 3), np.int32)
    lib().orc_kernel_map(_p(ic), ic.shape[0], _p(oc), oc.shape[0], int(ksize), int(dil), _p(nbr))
    return nbr


def pairs_from_nbr(nbr):
    """Canonical per-offset pair lists sorted by (k, out row): (pair_in, pair_out, k_offsets)."""
    K = nbr.shape[1]
    ins, outs, offs = [], [], [0]
    for k in range(K):
        o = np.nonzero(nbr[:, k] >= 0)[0]
        ins.append(nbr[o, k])
        outs.append(o)
        offs.append(offs[-1] + o.size)
    return (np.concatenate(ins).astype(np.int32), np.concatenate(outs).astype(np.int32),
            np.asarray(offs, np.int32))


def conv_fwd(feats, W, nbr, n_out):
    feats, W = _c(feats, np.float32), _c(W, np.float32)
    if W.ndim == 2:
        W = W[None]
    K, cin, cout = W.shape
    out = np.zeros((n_out, cout), np.float32)
    lib().orc_conv_fwd(_p(feats), _p(W), _p(_c(nbr, np.int32)), _p(out), n_out, K, cin, cout)
    return out


def conv_bwd(feats, W, gout, nbr):
    feats, W, gout = _c(feats, np.float32), _c(W, np.float32), _c(gout, np.float32)
    shape = W.shape
    if W.ndim == 2:
        W = W[None]
    K, cin, cout = W.shape
    gin = np.zeros_like(feats)
    gW = np.zeros_like(W)
    lib().orc_conv_bwd(_p(feats), _p(W), _p(gout), _p(_c(nbr, np.int32)), _p(gin), _p(gW), feats.shape[0],
                       gout.shape[0], K, cin, cout)
    return gin, gW.reshape(shape)


def convT_fwd(feats_coarse, W, nbr_down, n_fine):
    feats_coarse, W, nbr_down = _c(feats_coarse, np.float32), _c(W, np.float32), _c(nbr_down, np.int32)
    K, cin, cout = W.shape
    out = np.zeros((n_fine, cout), np.float32)
    lib().orc_convT_fwd(_p(feats_coarse), _p(W), _p(nbr_down), _p(out), n_fine, feats_coarse.shape[0], K, cin, cout)
    return out


def clusters_voxelize(clusters_idx, clusters_offset, coords, scale, spatial_shape, rand):
    """general_model.py:152-182: integer voxel coordinates [sumNPoint, 4] = (cluster id, x, y, z)."""
    idx, offs = _c(clusters_idx, np.int64), _c(clusters_offset, np.int32)
    coords, rand = _c(coords, np.float32), _c(rand, np.float32)
    out = np.zeros((idx.shape[0], 4), np.int32)
    lib().orc_clusters_voxelize(_p(idx), _p(offs), idx.shape[0], offs.shape[0] - 1, _p(coords),
                                ctypes.c_float(scale), int(spatial_shape), _p(rand), _p(out))
    return out


def tile_order(nbr):
    """Mask-sorted tile order of a 3^3 neighbour table (numpy restatement of csrc/tile_order.cu; builder-defined
    schedule, no reference counterpart): stable sort of the rows by
    key = [faces -x,+x,-y,+y,-z,+z of the stencil that hold a neighbour] << 26 | [27-bit mask without the centre].
    Returns (row_perm i32 [n], nbr_sorted i32 [n,27], tile_mask u32 [ceil(n/128)])."""
    nbr = np.ascontiguousarray(nbr, np.int32)
    n = nbr.shape[0]
    has = nbr >= 0
    k = np.arange(27)
    axes = (k % 3, (k // 3) % 3, k // 9)
    six = np.zeros(n, np.uint32)
    bit = 0
    for a in range(3):
        for side in (0, 2):
            six |= has[:, axes[a] == side].any(1).astype(np.uint32) << np.uint32(bit)
            bit += 1
    m = (has.astype(np.uint32) << k.astype(np.uint32)[None, :]).sum(1, dtype=np.uint32)
    m26 = (m & np.uint32(0x1FFF)) | ((m >> np.uint32(14)) << np.uint32(13))
    key = (six << np.uint32(26)) | m26
    perm = np.argsort(key, kind="stable").astype(np.int32)
    nbr_sorted = nbr[perm]
    tiles = (n + 127) // 128
    padded = np.zeros((tiles * 128, 27), bool)
    padded[:n] = nbr_sorted >= 0
    tile_has = padded.reshape(tiles, 128, 27).any(1)
    tile_mask = (tile_has.astype(np.uint32) << k.astype(np.uint32)[None, :]).sum(1, dtype=np.uint32)
    return perm, nbr_sorted, tile_mask
