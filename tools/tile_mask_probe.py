"""CPU probe behind the mask-sorted tile schedule (csrc/tile_order.cu): how many of the 27 offsets of a 3^3 kernel map
are active per 128-row tile under different row orders, on the benchmark's synthetic scenes (numpy only, no GPU).

    python tools/tile_mask_probe.py            # 4-scene benchmark batch, level 0

Round-1 result (M = 325,421 rows, rho = 9.11): first-occurrence order 26.96, Morton order 26.97 (sensor noise makes
surfaces two voxels thick), sort by the 27-bit mask 15.82, sort by (six stencil faces, mask) 12.68 -- the key the
CUDA kernel uses; variants with edge groups / face centres / entropy-ordered bits were within 0.3 of it.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minsu3d_b200.harness.scenes import make_scene  # noqa: E402

OFFS = np.array([(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)])


def _key(a):
    return (a[:, 0].astype(np.int64) + 2) * (1 << 40) + (a[:, 1].astype(np.int64) + 2) * (1 << 20) + (a[:, 2].astype(np.int64) + 2)


def scene_voxels(seed):
    xyz = make_scene(seed)["xyz"]
    v = np.floor((xyz - xyz.min(0)) / 0.02).astype(np.int64)
    _, first = np.unique(_key(v), return_index=True)
    return v[np.sort(first)]  # first-occurrence order (MinkowskiEngine's row order)


def neighbour_mask(c):
    sk = np.sort(_key(c))
    m = np.zeros((c.shape[0], 27), bool)
    for k, d in enumerate(OFFS):
        q = _key(c + d)
        pos = np.minimum(np.searchsorted(sk, q), c.shape[0] - 1)
        m[:, k] = sk[pos] == q
    return m


def morton(c):
    def part(v):
        v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
        for shift, mask in ((32, 0x1F00000000FFFF), (16, 0x1F0000FF0000FF), (8, 0x100F00F00F00F00F),
                            (4, 0x10C30C30C30C30C3), (2, 0x1249249249249249)):
            v = (v | (v << np.uint64(shift))) & np.uint64(mask)
        return v
    return part(c[:, 0]) | (part(c[:, 1]) << np.uint64(1)) | (part(c[:, 2]) << np.uint64(2))


def active_per_tile(nbr, perm, tile=128):
    n = nbr.shape[0]
    t = (n + tile - 1) // tile
    nb = np.concatenate([nbr[perm], np.zeros((t * tile - n, 27), bool)]).reshape(t, tile, 27)
    return nb.any(1).sum(1).mean()


def main():
    voxels = [scene_voxels(s) for s in range(4)]
    nbr = np.concatenate([neighbour_mask(c) for c in voxels])
    n = nbr.shape[0]
    w = (1 << np.arange(27)).astype(np.int64)
    mask = nbr.astype(np.int64) @ w
    six = np.stack([nbr[:, OFFS[:, a] == s].any(1) for a in range(3) for s in (-1, 1)], 1).astype(np.int64) @ (1 << np.arange(6))
    mort = np.concatenate([np.argsort(morton(c), kind="stable") + off for c, off in
                           zip(voxels, np.cumsum([0] + [len(c) for c in voxels[:-1]]))])
    print("rows %d  rho %.2f" % (n, nbr.sum() / n))
    print("first-occurrence order   %.2f" % active_per_tile(nbr, np.arange(n)))
    print("morton order             %.2f" % active_per_tile(nbr, mort))
    print("sort by mask             %.2f" % active_per_tile(nbr, np.argsort(mask, kind="stable")))
    print("sort by (six faces, mask) %.2f   <- tile_order.cu" % active_per_tile(nbr, np.lexsort((mask, six))))


if __name__ == "__main__":
    main()
