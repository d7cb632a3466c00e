"""Shared input generators for the parity tests (seeded, size-parameterised)."""
import numpy as np


def random_voxels(rng, n, extent=40, batch=2, dup=0.0):
    """int32 [n,4] coordinates; dup > 0 adds duplicated rows (for unique / quantize tests)."""
    c = np.concatenate((rng.integers(0, batch, (n, 1)), rng.integers(-extent // 4, extent, (n, 3))), 1).astype(np.int32)
    if dup > 0:
        k = int(n * dup)
        c[rng.integers(0, n, k)] = c[rng.integers(0, n, k)]
    return c


def surface_voxels(rng, n, batch=2):
    """Voxels on a few planes: ScanNet-like neighbourhood density (rho ~ 9)."""
    side = int(np.sqrt(n / (3 * batch))) + 1
    out = []
    for b in range(batch):
        u, v = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
        u, v = u.reshape(-1), v.reshape(-1)
        z = np.zeros_like(u)
        for (x, y, zz) in ((u, v, z), (u, z + 3, v), (z - 2, u, v)):
            out.append(np.stack((np.full_like(u, b), x, y, zz), 1))
    c = np.unique(np.concatenate(out).astype(np.int32), axis=0)
    c = c[rng.permutation(c.shape[0])]
    return np.ascontiguousarray(c[:n])


def clustered_points(rng, n, n_obj=12, batch=2, spread=0.05, collapse=False):
    """Points around object centres; collapse=True mimics shifted coordinates (dense blobs > 1000 nbrs)."""
    per = n // (n_obj * batch)
    xyz, lab, bidx = [], [], []
    for b in range(batch):
        for o in range(n_obj):
            c = rng.uniform(-2, 2, 3)
            s = 0.004 if collapse else spread
            xyz.append(c + rng.normal(0, s, (per, 3)))
            lab.append(np.full(per, 2 + (o % 5), np.int16))
            bidx.append(np.full(per, b, np.uint8))
    xyz = np.concatenate(xyz).astype(np.float32)
    lab, bidx = np.concatenate(lab), np.concatenate(bidx)
    # shuffle inside each batch so that indices are not grouped by object
    order = np.concatenate([np.nonzero(bidx == b)[0][rng.permutation((bidx == b).sum())] for b in range(batch)])
    xyz, lab, bidx = xyz[order], lab[order], bidx[order]
    offs = np.concatenate(([0], np.cumsum(np.bincount(bidx, minlength=batch)))).astype(np.int32)
    return xyz, lab, bidx, offs


def canon_clusters(cluster_idxs, cluster_offsets):
    """Order-independent view: list of sorted point arrays in cluster order."""
    return [np.sort(cluster_idxs[cluster_offsets[i]:cluster_offsets[i + 1], 1]) for i in range(len(cluster_offsets) - 1)]


def sg_mask_scores(seed, n_pairs, n_cls=19):
    """SoftGroup point-mask scores [S, classes+1] of the post-processing fixtures: regenerated from the seed on both
    sides (tests/golden/make_postproc_golden.py and the tests) instead of being stored (3 MB)."""
    rng = np.random.default_rng(10_000 + seed)
    return rng.normal(0.3, 1.0, (n_pairs, n_cls)).astype(np.float32)
