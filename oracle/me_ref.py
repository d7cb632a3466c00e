"""Dictionary-based restatement of the MinkowskiEngine semantics minsu3d relies on.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure Python / numpy, small cases only; written
to be obviously correct rather than fast.  PARITY UNPINNED against the real MinkowskiEngine binary
(un-vendored dependency, README.md:44-46,73); follows the reference call sites
minsu3d/model/module/common.py:7-95, backbone.py:8-43, general_model.py:152-193,
data/dataset/general_dataset.py:159-163 and SURVEY.md appendix A.
"""
import numpy as np


def sparse_quantize(coords, quantization_size=None):
    """appendix A.13: floor, first-occurrence unique -> (unique_map, inverse_map)."""
    c = np.asarray(coords)
    if quantization_size is not None:
        c = np.floor(c / quantization_size)
    c = np.floor(c).astype(np.int64)
    seen, unique_map, inverse = {}, [], []
    for r, row in enumerate(map(tuple, c)):
        if row not in seen:
            seen[row] = len(unique_map)
            unique_map.append(r)
        inverse.append(seen[row])
    return np.asarray(unique_map, np.int64), np.asarray(inverse, np.int64)


def stride_coords(coords, new_stride):
    """appendix A.3: out = floor(c / new_stride) * new_stride, unique in first-occurrence order."""
    c = np.asarray(coords, np.int64).copy()
    c[:, 1:] = np.floor_divide(c[:, 1:], new_stride) * new_stride
    um, inv = sparse_quantize(c)
    return c[um].astype(np.int32), inv


def kernel_offsets(ksize, dil):
    """appendix A.4: x fastest; odd: centred; even: 0..ksize-1."""
    lo = (ksize - 1) // 2 if ksize % 2 == 1 else 0
    offs = []
    for iz in range(ksize):
        for iy in range(ksize):
            for ix in range(ksize):
                offs.append(((ix - lo) * dil, (iy - lo) * dil, (iz - lo) * dil))
    return offs


def kernel_map(in_coords, out_coords, ksize, dil):
    """appendix A.5: list k holds (in_row, out_row) iff in_coord == out_coord + offset_k."""
    table = {tuple(int(v) for v in row): r for r, row in enumerate(np.asarray(in_coords))}
    maps = []
    for (dx, dy, dz) in kernel_offsets(ksize, dil):
        pairs = []
        for o, (b, x, y, z) in enumerate(np.asarray(out_coords)):
            i = table.get((int(b), int(x) + dx, int(y) + dy, int(z) + dz))
            if i is not None:
                pairs.append((i, o))
        maps.append(pairs)
    return maps


def conv_forward(feats, kernel, maps, n_out):
    """appendix A.7 in float64: out[o] += in[i] @ kernel[k], k ascending."""
    out = np.zeros((n_out, kernel.shape[-1]), np.float64)
    for k, pairs in enumerate(maps):
        for i, o in pairs:
            out[o] += feats[i].astype(np.float64) @ kernel[k].astype(np.float64)
    return out


def conv_transpose_forward(feats_coarse, kernel, maps_down, n_fine):
    """appendix A.6: forward strided map (fine i -> coarse o) with in/out swapped, same k."""
    out = np.zeros((n_fine, kernel.shape[-1]), np.float64)
    for k, pairs in enumerate(maps_down):
        for i, o in pairs:
            out[i] += feats_coarse[o].astype(np.float64) @ kernel[k].astype(np.float64)
    return out
