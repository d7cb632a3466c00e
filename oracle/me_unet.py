"""CPU restatement of the MinkUNet backbone forward on the MinkowskiEngine CPU-backend algorithm.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as the model-level
checker and by bench.py as the timed CPU baseline ("restatement of the MinkowskiEngine CPU backend
-- ME itself is an un-vendored dependency and is not installable"; PARITY UNPINNED vs the binary).

Follows minsu3d/model/module/backbone.py:8-43 and common.py:21-95 layer by layer, reading the
weights from a state dict with the reference's key names.  Sparse ops: oracle.c (hash kernel map,
per-offset gather-GEMM-scatter, OpenMP); BatchNorm/ReLU/Linear: numpy (float32, batch statistics).
"""
import numpy as np

from . import coord_unique, conv_fwd, convT_fwd, kernel_map


class _Maps:
    """Coordinate manager: coordinate maps per tensor stride + cached kernel maps (appendix A.15)."""

    def __init__(self, coords):
        self.coords = {1: np.ascontiguousarray(coords, np.int32)}
        self.k3 = {}
        self.k2 = {}

    def same(self, ts):
        if ts not in self.k3:
            c = self.coords[ts]
            self.k3[ts] = kernel_map(c, c, 3, ts)
        return self.k3[ts]

    def down(self, ts):
        if ts not in self.k2:
            fine = self.coords[ts]
            _, _, coarse = coord_unique(fine, 2 * ts)
            self.coords[2 * ts] = coarse
            self.k2[ts] = kernel_map(fine, coarse, 2, ts)
        return self.k2[ts]


def _bn(x, sd, prefix, training=True, eps=1e-5):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        mean = x.mean(0, dtype=np.float64)
        var = x.var(0, dtype=np.float64)
    else:
        mean, var = sd[prefix + ".running_mean"].astype(np.float64), sd[prefix + ".running_var"].astype(np.float64)
    y = (x - mean) / np.sqrt(var + eps) * w + b
    return y.astype(np.float32)


def _relu(x):
    return np.maximum(x, 0)


def _residual(x, sd, p, maps, ts, training):
    cin = x.shape[1]
    w1 = sd[p + ".conv_branch.2.kernel"]
    shortcut = x
    if (p + ".downsample.0.kernel") in sd:
        shortcut = x @ sd[p + ".downsample.0.kernel"]
    nbr = maps.same(ts)
    y = _relu(_bn(x, sd, p + ".conv_branch.0.bn", training))
    y = conv_fwd(y, w1, nbr, x.shape[0])
    y = _relu(_bn(y, sd, p + ".conv_branch.3.bn", training))
    y = conv_fwd(y, sd[p + ".conv_branch.5.kernel"], nbr, x.shape[0])
    assert cin == w1.shape[1]
    return y + shortcut


def _ublock(x, sd, p, maps, ts, depth, reps, training):
    for i in range(reps):
        x = _residual(x, sd, "%s.blocks.block%d" % (p, i), maps, ts, training)
    if depth > 1:
        skip = x
        nbr_down = maps.down(ts)
        n_coarse = maps.coords[2 * ts].shape[0]
        y = _relu(_bn(x, sd, p + ".conv.0.bn", training))
        y = conv_fwd(y, sd[p + ".conv.2.kernel"], nbr_down, n_coarse)
        y = _ublock(y, sd, p + ".u", maps, 2 * ts, depth - 1, reps, training)
        y = _relu(_bn(y, sd, p + ".deconv.0.bn", training))
        y = convT_fwd(y, sd[p + ".deconv.2.kernel"], nbr_down, x.shape[0])
        x = np.concatenate((skip, y), axis=1)
        for i in range(reps):
            x = _residual(x, sd, "%s.blocks_tail.block%d" % (p, i), maps, ts, training)
    return x


def unet_forward(state_dict, voxel_features, voxel_coords, prefix="backbone.unet", depth=7, reps=2,
                 training=True, has_input_conv=True):
    """Features after the U-Net's final BN+ReLU.  state_dict: name -> numpy array."""
    sd = state_dict
    maps = _Maps(voxel_coords)
    x = np.ascontiguousarray(voxel_features, np.float32)
    if has_input_conv:
        x = conv_fwd(x, sd[prefix + ".0.kernel"], maps.same(1), x.shape[0])
        ub, bn = prefix + ".1", prefix + ".2.bn"
    else:
        ub, bn = prefix + ".0", prefix + ".1.bn"
    x = _ublock(x, sd, ub, maps, 1, depth, reps, training)
    return _relu(_bn(x, sd, bn, training))


def _head(x, sd, p, training):
    y = x @ sd[p + ".0.weight"].T + sd[p + ".0.bias"]
    y = _relu(_bn(y, sd, p + ".1", training))
    return y @ sd[p + ".3.weight"].T + sd[p + ".3.bias"]


def backbone_forward(state_dict, voxel_features, voxel_coords, v2p_map, depth=7, reps=2, training=True):
    """backbone.py:36-43 -> dict(point_features, semantic_scores, point_offsets)."""
    feats = unet_forward(state_dict, voxel_features, voxel_coords, "backbone.unet", depth, reps, training)
    pf = feats[np.asarray(v2p_map)]
    return {"point_features": pf,
            "semantic_scores": _head(pf, state_dict, "backbone.semantic_branch", training),
            "point_offsets": _head(pf, state_dict, "backbone.offset_branch", training)}


def numpy_state_dict(module):
    return {k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}
