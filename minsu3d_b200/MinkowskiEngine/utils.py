"""ME.utils.sparse_quantize / sparse_collate / batched_coordinates (V1, V3).

Call sites: minsu3d/data/dataset/general_dataset.py:159-163 (numpy, DataLoader workers,
device="cpu"), minsu3d/model/general_model.py:187-189 (torch int64 [S,4], device="cuda"),
minsu3d/data/data_module.py:94-96 (sparse_collate).

Semantics: dc = floor(coords / quantization_size).int(); unique rows in
first-occurrence order; returns (dc[unique], feats[unique], unique_map, inverse_map) with int64 maps.

device="cuda" runs the libb2s coordinate hash.  device="cpu" is the host routine the reference's
DataLoader workers need (fork-safe, CUDA-free); it is an explicit API choice of the caller, never a
fallback: a CUDA request that cannot run raises.
"""
import numpy as np
import torch

from .. import ops


def _first_occurrence_unique_host(dc):
    """numpy restatement for the DataLoader-worker API (device='cpu')."""
    dc = np.ascontiguousarray(dc)
    _, first, inv = np.unique(dc, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique rows in first-occurrence order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return first[order].astype(np.int64), rank[inv.reshape(-1)].astype(np.int64)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    if labels is not None:
        raise NotImplementedError("label voting is not used by minsu3d and not supported")
    is_numpy = isinstance(coordinates, np.ndarray)
    device = str(device)
    if device.startswith("cuda"):
        coords_t = torch.as_tensor(coordinates)
        if not coords_t.is_cuda:
            coords_t = coords_t.to(device)
        if quantization_size is not None:
            coords_t = torch.floor(coords_t / quantization_size)
        elif coords_t.is_floating_point():
            coords_t = torch.floor(coords_t)
        dc = coords_t.to(torch.int32)
        n, d = dc.shape
        if d == 4:
            key_coords = dc.contiguous()
        elif d == 3:  # no batch column: pad a zero batch index for the hash
            key_coords = torch.cat((torch.zeros((n, 1), dtype=torch.int32, device=dc.device), dc), dim=1).contiguous()
        else:
            raise ValueError("coordinates must be [N,3] or [N,4]")
        _, unique_idx, inverse, _ = ops.coord_unique(key_coords, quant=1)
        unique_map = unique_idx.long()
        inverse_map = inverse.long()
        take = lambda t: t[unique_map]  # noqa: E731
        feats = None if features is None else torch.as_tensor(features).to(dc.device)
    elif device == "cpu":
        c = coordinates.detach().cpu().numpy() if torch.is_tensor(coordinates) else np.asarray(coordinates)
        if quantization_size is not None:
            c = np.floor(c / quantization_size)
        elif np.issubdtype(c.dtype, np.floating):
            c = np.floor(c)
        dc_np = c.astype(np.int32)
        um, im = _first_occurrence_unique_host(dc_np)
        unique_map, inverse_map = torch.from_numpy(um), torch.from_numpy(im)
        if is_numpy:
            dc = dc_np
            take = lambda t: t[um]  # noqa: E731
            feats = features
        else:
            dc = torch.from_numpy(dc_np)
            take = lambda t: t[unique_map]  # noqa: E731
            feats = features
    else:
        raise ValueError("unknown device %r" % device)

    if return_maps_only:
        return (unique_map, inverse_map) if return_inverse else unique_map
    ret = [take(dc)]
    if feats is not None:
        ret.append(take(feats))
    if return_index:
        ret.append(unique_map)
    if return_inverse:
        ret.append(inverse_map)
    return ret[0] if len(ret) == 1 else tuple(ret)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """Prepend the list position as batch index and concatenate (appendix A.14)."""
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c)
        if c.is_floating_point():
            c = torch.floor(c)
        c = c.to(dtype)
        out.append(torch.cat((torch.full((c.size(0), 1), b, dtype=dtype, device=c.device), c), dim=1))
    res = torch.cat(out, dim=0) if out else torch.zeros((0, 4), dtype=dtype)
    return res.to(device) if device is not None else res


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    bcoords = batched_coordinates(coords, dtype=dtype, device=device)
    bfeats = torch.cat([torch.as_tensor(f) for f in feats], dim=0)
    if device is not None:
        bfeats = bfeats.to(device)
    if labels is not None:
        return bcoords, bfeats, torch.cat([torch.as_tensor(l) for l in labels], dim=0)
    return bcoords, bfeats
