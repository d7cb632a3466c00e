"""ctypes binding of libb2s.so (include/b2s.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, the error is
raised.  torch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb2s.so")

_lib = None

c_void_p = ctypes.c_void_p
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float
c_size = ctypes.c_size_t

# name -> (restype, argtypes)
_P = c_void_p
_SIGNATURES = {
    "b2s_version": (c_i32, []),
    "b2s_last_error_string": (ctypes.c_char_p, []),
    "b2s_hash_capacity": (c_i64, [c_i64]),
    "b2s_coord_unique_ws_bytes": (c_size, [c_i64]),
    "b2s_coord_unique": (c_i32, [_P, c_i64, c_i32, _P, _P, c_i64, _P, _P, _P, _P, _P, _P, c_size, _P]),
    "b2s_kernel_map": (c_i32, [_P, c_i64, c_i32, c_i32, _P, _P, c_i64, _P, _P, _P]),
    "b2s_pairs_ws_bytes": (c_size, [c_i64, c_i32]),
    "b2s_pairs_from_nbr": (c_i32, [_P, c_i64, c_i32, c_i64, _P, _P, _P, _P, _P, c_size, _P]),
    "b2s_conv_ws_bytes": (c_size, [c_i32, c_i32, c_i32]),
    "b2s_conv_packed_floats": (c_i64, [c_i32, c_i32, c_i32]),
    "b2s_conv_pack": (c_i32, [_P, _P, c_i32, c_i32, c_i32, _P]),
    "b2s_conv_pack_multi": (c_i32, [_P, c_i32, c_i64, _P]),
    "b2s_conv_table": (c_i32, [_P, _P, _P, _P, _P, _P, _P, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, _P, c_size,
                               _P]),
    "b2s_conv_table_rows": (c_i32, [_P, _P, _P, _P, _P, _P, _P, _P, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, _P,
                                    c_size, _P]),
    "b2s_tile_order_ws_bytes": (c_size, [c_i64]),
    "b2s_tile_order": (c_i32, [_P, c_i64, c_i32, _P, _P, _P, _P, c_size, _P]),
    "b2s_conv_pairs": (c_i32, [_P, _P, _P, _P, _P, _P, _P, c_i32, c_i32, c_i32, c_i32, c_i64, c_i32, _P, c_size, _P]),
    "b2s_conv_wgrad": (c_i32, [_P, _P, _P, _P, _P, _P, c_i32, c_i32, c_i32, c_i64, c_i32, _P]),
    "b2s_conv_wgrad_ws_bytes": (c_size, [c_i32, c_i32, c_i32]),
    "b2s_conv_wgrad_ws": (c_i32, [_P, _P, _P, _P, _P, _P, c_i32, c_i32, c_i32, c_i64, c_i32, _P, c_size, _P]),
    "b2s_bn_ws_bytes": (c_size, [c_i64, c_i32]),
    "b2s_bn_stats": (c_i32, [_P, c_i64, c_i32, c_f32, c_f32, _P, _P, _P, _P, _P, _P, _P, c_size, _P]),
    "b2s_bn_forward": (c_i32, [_P, c_i64, c_i32, c_f32, c_f32, _P, _P, _P, _P, c_i32, _P, _P, _P, _P, _P, c_size, _P]),
    "b2s_bn_apply": (c_i32, [_P, c_i64, c_i32, _P, _P, _P, _P, c_i32, _P, _P]),
    "b2s_bn_backward": (c_i32, [_P, _P, _P, c_i64, c_i32, _P, _P, _P, c_i32, c_i32, _P, _P, _P, _P, _P, c_size, _P]),
    "b2s_bn_backward_add": (c_i32, [_P, _P, _P, _P, c_i64, c_i32, _P, _P, _P, c_i32, c_i32, _P, _P, _P, _P, _P, c_size,
                                    _P]),
    "b2s_resblock_ws_bytes": (c_size, [c_i32, c_i32, c_i32]),
    "b2s_resblock_forward": (c_i32, [_P, c_i64, c_i32, c_i32] + [_P] * 14 + [c_f32] * 4 + [_P] * 5 + [c_i32] + [_P] * 8 +
                             [c_i32, _P, c_size, _P]),
    "b2s_resblock_backward": (c_i32, [_P] * 15 + [c_i64, c_i32, c_i32] + [_P] * 5 + [c_i32] + [_P] * 3 + [c_i64] +
                              [_P] * 12 + [c_i32, _P, c_size, _P]),
    "b2s_bnconv_forward": (c_i32, [_P, c_i64, c_i32, c_i32, _P, _P, _P, _P, c_f32, c_f32, _P, _P, c_i32, _P, _P, _P, _P,
                                   _P, c_i64, c_i64, c_i64, c_i32, _P, _P, _P, _P, c_i32, _P, c_size, _P]),
    "b2s_bnconv_backward": (c_i32, [_P, _P, _P, _P, _P, _P, _P, c_i64, c_i32, c_i32, c_i32, _P, _P, _P, _P, _P, c_i64,
                                    c_i64, c_i64, c_i32, _P, _P, _P, _P, _P, c_i32, _P, c_size, _P]),
    "b2s_gather_rows": (c_i32, [_P, _P, c_i64, c_i32, _P, _P]),
    "b2s_scatter_add_rows": (c_i32, [_P, _P, c_i64, c_i32, _P, _P]),
    "b2s_ballquery_ws_bytes": (c_size, [c_i64]),
    "b2s_ballquery_count": (c_i32, [_P, _P, _P, c_i64, c_i32, c_f32, _P, _P, _P, c_size, _P]),
    "b2s_ballquery_fill": (c_i32, [_P, _P, _P, c_i64, c_i32, c_f32, _P, _P, _P, c_size, _P]),
    "b2s_cluster_ws_bytes": (c_size, [c_i64]),
    "b2s_cluster_label": (c_i32, [_P, _P, _P, c_i64, _P, _P, c_size, _P]),
    "b2s_cluster_select": (c_i32, [_P, _P, c_i64, c_i32, c_i32, c_f32, _P, c_i32, _P, _P, _P, _P, c_size, _P]),
    "b2s_cluster_order": (c_i32, [_P, _P, _P, _P, c_i64, c_i64, _P, _P, c_i32, _P, _P, c_size, _P]),
    "b2s_cluster_centers": (c_i32, [_P, _P, c_i32, _P, _P, _P, _P, _P]),
    "b2s_ha_assign": (c_i32, [_P, c_i32, _P, _P, c_i32, _P, _P, _P]),
    "b2s_ha_concat_ws_bytes": (c_size, [c_i32, c_i32]),
    "b2s_ha_concat": (c_i32, [_P, _P, c_i32, _P, _P, c_i32, _P, _P, _P, _P, c_size, _P]),
    "b2s_sec_mean": (c_i32, [_P, _P, _P, c_i32, c_i32, _P]),
    "b2s_sec_min": (c_i32, [_P, _P, _P, c_i32, c_i32, _P]),
    "b2s_sec_max": (c_i32, [_P, _P, _P, c_i32, c_i32, _P]),
    "b2s_roipool_fp": (c_i32, [_P, _P, _P, _P, c_i32, c_i32, _P]),
    "b2s_roipool_ws_bytes": (c_size, [c_i32, c_i32]),
    "b2s_roipool_fp_ws": (c_i32, [_P, _P, _P, _P, c_i32, c_i32, c_i64, _P, c_size, _P]),
    "b2s_roipool_bp": (c_i32, [_P, _P, _P, _P, c_i32, c_i32, _P]),
    "b2s_global_avg_pool_fp": (c_i32, [_P, _P, _P, c_i32, c_i32, _P]),
    "b2s_global_avg_pool_bp": (c_i32, [_P, _P, _P, c_i32, c_i32, _P]),
    "b2s_clusters_voxelize": (c_i32, [_P, c_i32, _P, c_i64, c_i32, _P, c_f32, c_i32, _P, _P, _P, _P]),
    "b2s_proposal_sort_ws_bytes": (c_size, [c_i64]),
    "b2s_proposal_sort": (c_i32, [_P, _P, c_i64, _P, _P, c_size, _P]),
    "b2s_proposal_npoint": (c_i32, [_P, c_i64, c_i32, _P, _P]),
    "b2s_proposal_iou": (c_i32, [_P, c_i64, _P, c_i32, _P, _P, _P]),
    "b2s_nms": (c_i32, [_P, _P, c_i32, c_f32, _P, _P, _P, c_size, _P]),
    "b2s_aug_affine": (c_i32, [_P, _P, c_i64, _P, _P, _P, _P, _P]),
    "b2s_elastic_blur": (c_i32, [_P, _P, c_i32, c_i32, c_i32, _P]),
    "b2s_elastic_apply": (c_i32, [_P, _P, c_i64, c_i32, c_i32, c_i32, ctypes.c_double, ctypes.c_double, _P]),
    "b2s_crop_test": (c_i32, [_P, c_i64, _P, _P, _P, _P, _P, _P]),
    "b2s_cross_entropy_ws_bytes": (c_size, [c_i64]),
    "b2s_cross_entropy_forward": (c_i32, [_P, _P, c_i32, c_i64, c_i32, c_i32, _P, _P, _P, _P, c_size, _P]),
    "b2s_cross_entropy_backward": (c_i32, [_P, _P, c_i32, c_i64, c_i32, c_i32, _P, _P, _P, _P]),
    "b2s_get_iou": (c_i32, [_P, _P, _P, _P, _P, _P, c_i32, c_i32, _P]),
    "b2s_get_mask_label": (c_i32, [_P, _P, _P, _P, _P, c_i32, c_i32, c_i32, c_f32, _P, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load libb2s.so once; raise loudly when it is absent (no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libb2s.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python minsu3d_b200/csrc/build.py`. There is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        ns = _Namespace()
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
            k = KERNELS_PER_CALL.get(name, 0)
            setattr(ns, name, _counted(fn, k) if k else fn)
        _lib = ns
    return _lib


class _Namespace:
    pass


# kernels launched by one call of each entry point (CUB scan = 2, 64-bit radix sort ~ 10);
# used for the `gpu_launches` figure of bench.py
KERNELS_PER_CALL = {
    "b2s_coord_unique": 6, "b2s_kernel_map": 1, "b2s_pairs_from_nbr": 4, "b2s_conv_pack": 1, "b2s_conv_pack_multi": 1, "b2s_conv_table": 1, "b2s_conv_table_rows": 1, "b2s_tile_order": 7, "b2s_conv_pairs": 1,
    "b2s_conv_wgrad": 1, "b2s_conv_wgrad_ws": 1, "b2s_resblock_forward": 4, "b2s_resblock_backward": 6, "b2s_bnconv_forward": 2, "b2s_bnconv_backward": 3, "b2s_bn_backward_add": 1, "b2s_bn_stats": 1, "b2s_bn_forward": 1, "b2s_bn_apply": 1, "b2s_bn_backward": 1, "b2s_gather_rows": 1,
    "b2s_scatter_add_rows": 1, "b2s_ballquery_count": 16, "b2s_ballquery_fill": 2, "b2s_cluster_label": 6,
    "b2s_cluster_select": 7, "b2s_cluster_order": 4, "b2s_cluster_centers": 1, "b2s_ha_assign": 1,
    "b2s_ha_concat": 4, "b2s_sec_mean": 1, "b2s_sec_min": 1, "b2s_sec_max": 1, "b2s_roipool_fp": 1, "b2s_roipool_fp_ws": 2,
    "b2s_roipool_bp": 1, "b2s_global_avg_pool_fp": 1, "b2s_global_avg_pool_bp": 1, "b2s_get_iou": 1, "b2s_clusters_voxelize": 2,
    "b2s_get_mask_label": 1, "b2s_cross_entropy_forward": 1, "b2s_cross_entropy_backward": 1, "b2s_aug_affine": 1, "b2s_elastic_blur": 6, "b2s_elastic_apply": 1, "b2s_crop_test": 1, "b2s_proposal_sort": 10, "b2s_proposal_npoint": 1, "b2s_proposal_iou": 2, "b2s_nms": 1,
}
_launches = [0]


def _counted(fn, k):
    def call(*a):
        _launches[0] += k
        return fn(*a)
    return call


def reset_launch_count():
    _launches[0] = 0


def launch_count():
    return _launches[0]


def check(rc, what=""):
    if rc != 0:
        msg = lib().b2s_last_error_string().decode()
        raise RuntimeError("libb2s %s failed with code %d: %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def stream():
    """Raw cudaStream_t of torch's current stream on the current device (fast path: ~0.3 us)."""
    if _raw_stream is not None:
        # torch._C._cuda_getDevice skips torch.cuda.current_device()'s lazy-init checks (~1.2 us per call, called
        # about twice per library call)
        return _raw_stream(_get_device())
    return torch.cuda.current_stream().cuda_stream


_WS = {}


def workspace(nbytes, device, slot=0):
    """Grow-only scratch buffer per (device, stream, slot); stream-ordered reuse is safe.  slot > 0: additional
    buffers for calls whose scratch must survive other library calls (batched ball queries / clusterings)."""
    key = (device.index, stream(), slot)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError("expected a CUDA tensor (libb2s has no CPU path)")


def require(cond, msg):
    if not cond:
        raise ValueError(msg)
