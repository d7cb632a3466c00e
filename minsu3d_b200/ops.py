"""Tensor-level wrappers over the C ABI (include/b2s.h).  torch = memory + streams only.

Every function validates dtype/device/contiguity (TypeError / ValueError, mirroring the
reference wrappers' asserts, minsu3d/common_ops/functions/common_ops.py:27-29) and then passes raw
device pointers and the current stream to libb2s.  No function here has a CPU branch.
"""
import torch

from . import _cabi
from ._cabi import check, lib, ptr, require, require_cuda, stream, workspace

I32 = torch.int32


def _dev_i32(n, device):
    return torch.empty(n, dtype=I32, device=device)


# ------------------------------------------------------------------------------------------
# T1 / V1 coordinate hash
# ------------------------------------------------------------------------------------------
class HashTable:
    """Open-addressing coordinate table living in torch-owned device memory."""

    __slots__ = ("keys", "vals", "cap")

    def __init__(self, n, device):
        self.cap = int(lib().b2s_hash_capacity(int(n)))
        self.keys = torch.empty(self.cap, dtype=torch.int64, device=device)
        self.vals = torch.empty(self.cap, dtype=I32, device=device)


def coord_unique_async(coords, quant=1, d_n=None):
    """Enqueue insert + unique without reading the count back.

    coords may be an upper-bound-sized buffer whose real row count is the device int32 d_n[0].
    Returns (table, unique_idx_buf, inverse_buf, out_coords_buf, d_count) -- all sized by the upper bound.
    """
    require_cuda(coords)
    require(coords.dtype == I32 and coords.dim() == 2 and coords.size(1) == 4 and coords.is_contiguous(),
            "coords must be a contiguous int32 [n,4] tensor")
    n = coords.size(0)
    dev = coords.device
    table = HashTable(n, dev)
    unique_idx = _dev_i32(n, dev)
    inverse = _dev_i32(n, dev)
    out_coords = torch.empty((n, 4), dtype=I32, device=dev)
    d_count = _dev_i32(2, dev)
    ws = workspace(lib().b2s_coord_unique_ws_bytes(n), dev)
    check(lib().b2s_coord_unique(ptr(coords), n, int(quant), ptr(table.keys), ptr(table.vals), table.cap,
                                 ptr(unique_idx), ptr(inverse), ptr(out_coords), ptr(d_count), ptr(d_n), ptr(ws),
                                 ws.numel(), stream()), "coord_unique")
    return table, unique_idx, inverse, out_coords, d_count


# Deferred validation of device-side counts.  Data-dependent sizes normally cost a host read (the GPU drains, the
# host loses its run-ahead); when the caller KNOWS the size (coordinates that are unique by construction, level
# sizes the voxeliser reported) the read is skipped, the device count is kept, and it is compared with the claim
# at the next host read that happens anyway.  A wrong claim raises there -- it is never silently accepted.
_RANGE_MSG = "coordinate outside the packable range (batch < 2^19, |xyz| < 2^14)"
_DEFERRED = []  # (d_count int32[2] = {m, range_error}, expected m, what)


def defer_count_check(d_count, expected, what):
    _DEFERRED.append((d_count, int(expected), what))
    if len(_DEFERRED) >= 64:  # nobody synchronised for a long time (a few backbone-only forwards): do it now
        run_deferred_checks()


def run_deferred_checks():
    """Validate the pending size claims; call right after a host read (the stream is drained, this is cheap)."""
    if not _DEFERRED:
        return
    items = _DEFERRED[:]
    del _DEFERRED[:]
    vals = torch.stack([d for d, _, _ in items]).tolist()
    for (m, bad), (_, expected, what) in zip(vals, items):
        if bad:
            raise ValueError(_RANGE_MSG)
        if m != expected:
            raise ValueError("%s: claimed %d rows but the device counted %d" % (what, expected, m))


def deferred_failure_flag(device):
    """float32 scalar device tensor: 1.0 iff any pending size claim is wrong -- computed on the device, NO host read.

    For loops that update weights before the next host read (Trainer.step): hand it to the fused optimizer as
    `found_inf`, which then skips the update on the device, so a wrong claim (rows sliced to a wrong count,
    un-subsampled features) can never reach the weights; the ValueError itself is raised by the next
    run_deferred_checks().  Returns None when nothing is pending."""
    if not _DEFERRED:
        return None
    counts = torch.stack([d for d, _, _ in _DEFERRED])
    expected = torch.tensor([e for _, e, _ in _DEFERRED], dtype=I32).to(device, non_blocking=True)
    bad = (counts[:, 0] != expected) | (counts[:, 1] != 0)
    return bad.any().to(torch.float32).reshape(())


def coord_unique(coords, quant=1, assume_unique=False):
    """First-occurrence unique of int32 [n,4] coordinates (optionally floor-quantised).

    Returns (table, unique_idx[m] i32, inverse[n] i32, out_coords[m,4] i32).  One host read of m -- unless
    assume_unique: the caller states that the rows are already unique (m == n), no host read happens and the
    statement is validated later (run_deferred_checks).
    """
    table, unique_idx, inverse, out_coords, d_count = coord_unique_async(coords, quant)
    if assume_unique:
        defer_count_check(d_count, coords.size(0), "SparseTensor(coordinates_unique=True)")
        return table, unique_idx, inverse, out_coords
    m, bad = d_count.tolist()
    run_deferred_checks()
    if bad:
        raise ValueError(_RANGE_MSG)
    return table, unique_idx[:m], inverse, out_coords[:m]


def coord_pyramid(coords, base_stride, levels, size_hints=None):
    """Strided maps base*2, base*4, ... built back to back on the device; ONE host read for all counts.

    size_hints: the row counts of the levels as reported by whoever voxelised the scene (a by-product of the data
    loader's quantisation); then there is NO host read, the device counts are validated later.
    Returns a list of (stride, table, out_coords[m_l, 4]).
    """
    out, d_counts = [], []
    cur, d_n, stride = coords, None, base_stride
    for _ in range(levels):
        stride *= 2
        table, _, _, oc, d_count = coord_unique_async(cur, stride, d_n)
        out.append((stride, table, oc))
        d_counts.append(d_count)
        cur, d_n = oc, d_count
    if size_hints is not None:
        require(len(size_hints) >= levels, "size_hints must cover every level")
        for (s, _, _), d_count, m in zip(out, d_counts, size_hints):
            defer_count_check(d_count, m, "coordinate map of tensor stride %d (size hint)" % s)
        return [(s, t, oc[:int(m)]) for (s, t, oc), m in zip(out, size_hints)]
    counts = torch.stack(d_counts).tolist()
    run_deferred_checks()
    if any(bad for _, bad in counts):
        raise ValueError(_RANGE_MSG)
    return [(s, t, oc[:m]) for (s, t, oc), (m, _) in zip(out, counts)]


def kernel_map(out_coords, table, ksize, dil, with_tile_mask=False):
    """Output-stationary neighbour table nbr[n_out, ksize^3] (int32, -1 = no input).

    with_tile_mask: also return uint32-as-int32 [ceil(n_out / 128)] active-offset masks (bit k set iff some row of
    the 128-row tile has a neighbour at offset k) that conv_table uses to skip empty offsets.
    """
    require_cuda(out_coords)
    require(out_coords.dtype == I32 and out_coords.is_contiguous(), "out_coords must be contiguous int32")
    n_out = out_coords.size(0)
    K = ksize ** 3
    nbr = torch.empty((n_out, K), dtype=I32, device=out_coords.device)
    tile_mask = None
    if with_tile_mask and K <= 32:
        tile_mask = torch.empty(((n_out + 127) // 128,), dtype=I32, device=out_coords.device)
    check(lib().b2s_kernel_map(ptr(out_coords), n_out, int(ksize), int(dil), ptr(table.keys), ptr(table.vals),
                               table.cap, ptr(nbr), ptr(tile_mask), stream()), "kernel_map")
    return (nbr, tile_mask) if with_tile_mask else nbr


def tile_order(nbr):
    """Mask-sorted tile order of a 3^3 neighbour table: (row_perm[n] i32, nbr_sorted[n,27] i32, tile_mask i32).

    A schedule for conv_table(..., out_rows=row_perm): tiles of 128 sorted rows share their neighbour pattern, so
    the tcgen05 kernel skips most empty offsets.  Results are unchanged.
    """
    require_cuda(nbr)
    require(nbr.dtype == I32 and nbr.dim() == 2 and nbr.size(1) == 27 and nbr.is_contiguous(),
            "nbr must be a contiguous int32 [n,27] table")
    n = nbr.size(0)
    dev = nbr.device
    row_perm = _dev_i32(n, dev)
    nbr_sorted = torch.empty_like(nbr)
    tile_mask = _dev_i32((n + 127) // 128, dev)
    ws = workspace(lib().b2s_tile_order_ws_bytes(n), dev)
    check(lib().b2s_tile_order(ptr(nbr), n, 27, ptr(row_perm), ptr(nbr_sorted), ptr(tile_mask), ptr(ws), ws.numel(),
                               stream()), "tile_order")
    return row_perm, nbr_sorted, tile_mask


def pairs_from_nbr(nbr, exact=False):
    """Canonical pair lists: (pair_in, pair_out, k_offsets[K+1], d_count[1]) sorted by (k, out row).

    Arrays are allocated at the n_out*K upper bound (no host sync); pass exact=True to trim.
    """
    n_out, K = nbr.shape
    dev = nbr.device
    cap = max(n_out * K, 1)
    pair_in = _dev_i32(cap, dev)
    pair_out = _dev_i32(cap, dev)
    k_offsets = _dev_i32(K + 1, dev)
    d_count = _dev_i32(1, dev)
    ws = workspace(lib().b2s_pairs_ws_bytes(n_out, K), dev)
    check(lib().b2s_pairs_from_nbr(ptr(nbr), n_out, K, cap, ptr(pair_in), ptr(pair_out), ptr(k_offsets),
                                   ptr(d_count), ptr(ws), ws.numel(), stream()), "pairs_from_nbr")
    if exact:
        p = int(d_count.item())
        return pair_in[:p], pair_out[:p], k_offsets, p
    return pair_in, pair_out, k_offsets, d_count


# ------------------------------------------------------------------------------------------
# T3 / T4 convolution products
# ------------------------------------------------------------------------------------------
ALGO_AUTO, ALGO_SIMT, ALGO_TC_3XTF32, ALGO_TC_TF32, ALGO_WARP_STREAM = 0, 1, 2, 3, 4
_default_algo = ALGO_AUTO


def set_conv_algo(algo):
    global _default_algo
    _default_algo = int(algo)


def get_conv_algo():
    return _default_algo


def _f32c(t, name):
    require_cuda(t)
    require(t.dtype == torch.float32 and t.is_contiguous(), "%s must be contiguous float32" % name)


_CONV_WS_BYTES = {}


def _conv_ws(K, c_in, c_out, device):
    key = (K, c_in, c_out)
    nbytes = _CONV_WS_BYTES.get(key)
    if nbytes is None:
        nbytes = _CONV_WS_BYTES[key] = lib().b2s_conv_ws_bytes(K, c_in, c_out)
    return workspace(nbytes, device)


def conv_tc_shape_ok(K, c_in, c_out):
    """Shapes the tcgen05 path takes (conv_tc_supported in csrc/conv_tc.cu)."""
    return 1 <= K <= 32 and c_in >= 16 and c_in % 16 == 0 and c_out >= 16 and c_out % 16 == 0 and c_out <= 256


def conv_packed_floats(K, c_in, c_out):
    """Floats of the packed image of W[K, c_in, c_out] (both orientations), 0 when the tcgen05 path does not take it."""
    if not (conv_tc_shape_ok(K, c_in, c_out) and conv_tc_shape_ok(K, c_out, c_in)):
        return 0
    return 4 * K * c_in * c_out


def conv_pack(W, out=None):
    """b2s_conv_pack: W [K, c_in, c_out] (or [c_in, c_out]) -> the tensor-core operand images, both orientations, one
    launch.  Pass the result as `packed=` to conv_table / conv_pairs while W is unchanged (once per optimizer step)."""
    _f32c(W, "W")
    K, c_in, c_out = (1,) + tuple(W.shape) if W.dim() == 2 else tuple(W.shape)
    n = conv_packed_floats(K, c_in, c_out)
    require(n > 0, "conv_pack: shape not taken by the tcgen05 path")
    if out is None or out.numel() != n:
        out = torch.empty(n, dtype=torch.float32, device=W.device)
    check(lib().b2s_conv_pack(ptr(W), ptr(out), K, c_in, c_out, stream()), "conv_pack")
    return out


# Every torch optimizer step invalidates the packed images: the fused (multi-tensor CUDA) optimizers update the
# parameters WITHOUT moving their version counters (torch.optim.Adam(fused=True): `_version` stays put), so the
# version alone cannot be trusted.  The hook is global (torch.optim.optimizer), cheap (one integer), and registered
# once at import.  Writers that bypass both mechanisms (raw kernels on `data_ptr()`) must call invalidate_packed().
_PACK_EPOCH = [0]


def invalidate_packed(*_args, **_kwargs):
    _PACK_EPOCH[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook
    _register_step_hook(invalidate_packed)
except ImportError:  # pragma: no cover -- torch < 2.0
    _register_step_hook = None


class PackedWeights:
    """Per-parameter cache of conv_pack(): re-packed when the parameter was modified in place (its torch version
    counter moved: load_state_dict, init, in-place ops), replaced or moved, or when ANY optimizer has stepped since
    (see _PACK_EPOCH).  Round 1 re-packed at every product: 170 launches per PointGroup step on the critical path
    of forward and backward."""

    __slots__ = ("ref", "version", "addr", "epoch", "buf")

    def __init__(self):
        self.ref, self.version, self.addr, self.epoch, self.buf = None, -1, 0, -1, None

    def __deepcopy__(self, memo):
        return PackedWeights()

    def get(self, W):
        if not W.is_cuda:
            return None
        if (self.ref is W and self.version == W._version and self.addr == W.data_ptr()
                and self.epoch == _PACK_EPOCH[0] and self.buf is not None):
            return self.buf
        K, c_in, c_out = (1,) + tuple(W.shape) if W.dim() == 2 else tuple(W.shape)
        if conv_packed_floats(K, c_in, c_out) == 0 or not W.is_contiguous():
            return None
        with torch.no_grad():
            reuse = self.buf if (self.buf is not None and self.buf.device == W.device) else None
            self.buf = conv_pack(W.detach(), reuse)
        self.ref, self.version, self.addr, self.epoch = W, W._version, W.data_ptr(), _PACK_EPOCH[0]
        return self.buf


class PackedSet:
    """All tensor-core-eligible convolution kernels of a model, re-packed in ONE launch (b2s_conv_pack_multi).

    `caches`: list of (parameter, PackedWeights) -- e.g. (conv.kernel, conv._packed) of every MinkowskiConvolution.
    repack() refreshes every image and stamps the per-parameter caches as current, so the modules' lazy
    `PackedWeights.get()` finds them valid: call it right after optimizer.step()."""

    def __init__(self, caches):
        self.items = []
        rows, start = [], 0
        for W, cache in caches:
            if not (W.is_cuda and W.is_contiguous() and W.dtype == torch.float32):
                continue
            K, c_in, c_out = (1,) + tuple(W.shape) if W.dim() == 2 else tuple(W.shape)
            n = conv_packed_floats(K, c_in, c_out)
            if n == 0:
                continue
            buf = torch.empty(n, dtype=torch.float32, device=W.device)
            rows.append([W.data_ptr(), buf.data_ptr(), K, c_in, c_out, start])
            start += K * c_in * c_out
            self.items.append((W, cache, buf))
        self.total = start
        self.desc = (torch.tensor(rows, dtype=torch.int64, device=self.items[0][0].device) if rows else None)

    def repack(self):
        if self.desc is None:
            return
        for W, _, buf in self.items:  # the table holds raw addresses: parameters must not have been moved
            require(W.data_ptr() == int(self._addr(W)), "PackedSet: a parameter was moved; rebuild the set")
        check(lib().b2s_conv_pack_multi(ptr(self.desc), len(self.items), self.total, stream()), "conv_pack_multi")
        for W, cache, buf in self.items:
            cache.ref, cache.version, cache.addr, cache.epoch, cache.buf = W, W._version, W.data_ptr(), _PACK_EPOCH[0], buf

    def _addr(self, W):
        if not hasattr(self, "_addr_of"):
            self._addr_of = {id(w): w.data_ptr() for w, _, _ in self.items}
        return self._addr_of[id(W)]


def conv_table(A, W, nbr, n_out, K, c_in, c_out, w_transposed=False, k_reversed=False, algo=None, tile_mask=None,
               out_rows=None, packed=None, add_src=None):
    """out_rows: nbr / tile_mask are the mask-sorted table of tile_order() and out_rows its row permutation.
    packed: conv_pack(W) (skips the per-call packing); add_src [n_out, c_out]: residual added to the result."""
    _f32c(A, "A")
    _f32c(W, "W")
    if add_src is not None:
        _f32c(add_src, "add_src")
        require(tuple(add_src.shape) == (n_out, c_out), "add_src must be [n_out, c_out]")
    out = torch.empty((n_out, c_out), dtype=torch.float32, device=A.device)
    ws = _conv_ws(K, c_in, c_out, A.device)
    if out_rows is not None:
        check(lib().b2s_conv_table_rows(ptr(A), ptr(W), ptr(packed), ptr(nbr), ptr(tile_mask), ptr(out_rows),
                                        ptr(add_src), ptr(out), n_out, K, c_in, c_out, int(w_transposed),
                                        int(k_reversed), _default_algo if algo is None else algo, ptr(ws), ws.numel(),
                                        stream()),
              "conv_table_rows")
        return out
    check(lib().b2s_conv_table(ptr(A), ptr(W), ptr(packed), ptr(nbr), ptr(tile_mask), ptr(add_src), ptr(out), n_out, K,
                               c_in, c_out, int(w_transposed), int(k_reversed),
                               _default_algo if algo is None else algo, ptr(ws), ws.numel(), stream()), "conv_table")
    return out


def conv_pairs(A, W, src, dst, k_offsets, n_out, K, c_in, c_out, max_pairs, w_transposed=False,
               zero_init=False, algo=None, packed=None):
    _f32c(A, "A")
    _f32c(W, "W")
    alloc = torch.zeros if zero_init else torch.empty
    out = alloc((n_out, c_out), dtype=torch.float32, device=A.device)
    ws = _conv_ws(K, c_in, c_out, A.device)
    check(lib().b2s_conv_pairs(ptr(A), ptr(W), ptr(packed), ptr(src), ptr(dst), ptr(k_offsets), ptr(out), K, c_in, c_out,
                               int(w_transposed), int(max_pairs), _default_algo if algo is None else algo,
                               ptr(ws), ws.numel(), stream()), "conv_pairs")
    return out


def conv_wgrad(A, G, src, dst, k_offsets, K, c_a, c_g, max_pairs, algo=None):
    _f32c(A, "A")
    _f32c(G, "G")
    gW = torch.empty((K, c_a, c_g), dtype=torch.float32, device=A.device)
    ws = workspace(wgrad_ws_bytes(K, c_a, c_g), A.device)
    check(lib().b2s_conv_wgrad_ws(ptr(A), ptr(G), ptr(src), ptr(dst), ptr(k_offsets), ptr(gW), K, c_a, c_g,
                                  int(max_pairs), _default_algo if algo is None else algo, ptr(ws), ws.numel(),
                                  stream()), "conv_wgrad")
    return gW


_WGRAD_WS = {}


def wgrad_ws_bytes(K, c_a, c_g):
    key = (K, c_a, c_g)
    v = _WGRAD_WS.get(key)
    if v is None:
        v = _WGRAD_WS[key] = lib().b2s_conv_wgrad_ws_bytes(K, c_a, c_g)
    return v


# ------------------------------------------------------------------------------------------
# Point-wise linear heads (backbone.py:21-35: nn.Linear over [n_points, m]).  Forward and data gradient stay
# torch/cuBLAS (library GEMMs); the WEIGHT gradient dW = dy^T x contracts over 400k points into a 16 x 20 matrix --
# cuBLAS picks a SIMT split-K kernel for that shape (0.36 ms for one head) -- and is the K = 1 identity-map case of
# the T3 weight gradient, so it runs on b2s_conv_wgrad_ws (deterministic, any channel count).
# ------------------------------------------------------------------------------------------
_IDENT = {}


def _ident_pairs(n, device):
    key = (n, device.index)
    v = _IDENT.get(key)
    if v is None:
        if len(_IDENT) > 64:
            _IDENT.clear()
        v = _IDENT[key] = (torch.arange(n, dtype=I32, device=device), torch.tensor([0, n], dtype=I32, device=device))
    return v


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gy @ weight
        if ctx.needs_input_grad[1]:
            n = x.size(0)
            ident, koff = _ident_pairs(n, x.device)
            gw = conv_wgrad(gy, x.contiguous(), ident, ident, koff, 1, weight.size(0), weight.size(1), n)[0]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb


LINEAR_MIN_ROWS = 16384


def linear(x, weight, bias=None):
    """F.linear with the weight gradient on libb2s for tall fp32 CUDA inputs; plain F.linear otherwise."""
    if (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.size(0) >= LINEAR_MIN_ROWS
            and torch.is_grad_enabled() and weight.requires_grad):
        return _LinearFn.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


# ------------------------------------------------------------------------------------------
# T5 batch norm
# ------------------------------------------------------------------------------------------
_BN_COUNTERS = {}


_RESBLOCK_WS = {}


def resblock_ws_bytes(K, c_in, c_out):
    key = (K, c_in, c_out)
    v = _RESBLOCK_WS.get(key)
    if v is None:
        v = _RESBLOCK_WS[key] = lib().b2s_resblock_ws_bytes(K, c_in, c_out)
    return v


def bn_counter(device):
    return _bn_counter(device)


def _bn_counter(device):
    """Per-(device, stream) int32 that is zero between launches (last-block-done second stage)."""
    key = (device.index, stream())
    t = _BN_COUNTERS.get(key)
    if t is None:
        t = _BN_COUNTERS[key] = torch.zeros(1, dtype=I32, device=device)
    return t


def bn_stats(x, eps=1e-5, momentum=0.0, running_mean=None, running_var=None):
    """Batch statistics in one call: returns (mean, rstd); updates the running statistics in place."""
    _f32c(x, "x")
    n, c = x.shape
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    rstd = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = _bn_ws(c, x.device)
    check(lib().b2s_bn_stats(ptr(x), n, c, float(eps), float(momentum), ptr(running_mean), ptr(running_var),
                             ptr(mean), None, ptr(rstd), ptr(_bn_counter(x.device)), ptr(ws), ws.numel(), stream()),
          "bn_stats")
    return mean, rstd


def _bn_ws(c, device):
    # b2s_bn_ws_bytes does not depend on n (the grid of the column-sum kernel is capped): 296 * 2 * c doubles
    return workspace(296 * 2 * max(int(c), 4) * 8 + 2048, device)


def bn_forward(x, eps, momentum, running_mean, running_var, gamma, beta, relu):
    """Training-mode BatchNorm(+ReLU) forward in one library call: returns (y, mean, rstd)."""
    _f32c(x, "x")
    n, c = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((2, c), dtype=torch.float32, device=x.device)
    mean, rstd = stats[0], stats[1]
    ws = _bn_ws(c, x.device)
    check(lib().b2s_bn_forward(ptr(x), n, c, float(eps), float(momentum), ptr(running_mean), ptr(running_var),
                               ptr(gamma), ptr(beta), int(relu), ptr(y), ptr(mean), ptr(rstd),
                               ptr(_bn_counter(x.device)), ptr(ws), ws.numel(), stream()), "bn_forward")
    return y, mean, rstd


def bn_apply(x, mean, rstd, gamma, beta, relu, out=None):
    _f32c(x, "x")
    n, c = x.shape
    y = torch.empty_like(x) if out is None else out
    check(lib().b2s_bn_apply(ptr(x), n, c, ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), int(relu), ptr(y),
                             stream()), "bn_apply")
    return y


def bn_backward(x, y, dy, mean, rstd, gamma, relu, training):
    _f32c(x, "x")
    _f32c(dy, "dy")
    n, c = x.shape
    dx = torch.empty_like(x)
    dgb = torch.empty((2, c), dtype=torch.float32, device=x.device)
    dgamma, dbeta = dgb[0], dgb[1]
    ws = _bn_ws(c, x.device)
    check(lib().b2s_bn_backward(ptr(x), ptr(y), ptr(dy), n, c, ptr(mean), ptr(rstd), ptr(gamma), int(relu),
                                int(training), ptr(dx), ptr(dgamma), ptr(dbeta), ptr(_bn_counter(x.device)), ptr(ws),
                                ws.numel(), stream()),
          "bn_backward")
    return dx, dgamma, dbeta


# ------------------------------------------------------------------------------------------
# V2 devoxelise
# ------------------------------------------------------------------------------------------
class _Devoxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, idx):
        _f32c(feat, "feat")
        require_cuda(idx)
        require(idx.dtype == torch.int64 and idx.is_contiguous(), "idx must be contiguous int64")
        n, c = idx.numel(), feat.size(1)
        out = torch.empty((n, c), dtype=torch.float32, device=feat.device)
        check(lib().b2s_gather_rows(ptr(feat), ptr(idx), n, c, ptr(out), stream()), "gather_rows")
        ctx.save_for_backward(idx)
        ctx.m = feat.size(0)
        return out

    @staticmethod
    def backward(ctx, grad):
        (idx,) = ctx.saved_tensors
        grad = grad.contiguous()
        n, c = grad.shape
        gfeat = torch.zeros((ctx.m, c), dtype=torch.float32, device=grad.device)
        check(lib().b2s_scatter_add_rows(ptr(grad), ptr(idx), n, c, ptr(gfeat), stream()), "scatter_add_rows")
        return gfeat, None


def devoxelize(feat, idx):
    """feat[idx] with a vectorised scatter-add gradient (backbone.py:40, pointgroup.py:88)."""
    return _Devoxelize.apply(feat, idx)


# ------------------------------------------------------------------------------------------
# C1 ball query
# ------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------
# fork / join over side streams: independent library call sequences (PointGroup's two ball queries and clusterings,
# pointgroup.py:43-66) run concurrently -- the BFS order of one (latency-bound: three device-wide barriers per BFS
# level) overlaps the union-find / list fill of the other (throughput-bound).  Outputs are allocated on the main
# stream before the fork and main waits for every side stream before anything is returned, so the caching
# allocator's stream-ordered reuse stays valid.  B2S_CLUSTER_STREAMS=0: everything on the current stream.
# ------------------------------------------------------------------------------------------
import contextlib
import os

_SIDE_STREAMS = {}
_FORK_ON = os.environ.get("B2S_CLUSTER_STREAMS", "1") != "0"


class _Fork:
    def __init__(self, device, n):
        self.main = torch.cuda.current_stream(device)
        key = (device.index, self.main.cuda_stream)
        pool = _SIDE_STREAMS.setdefault(key, [])
        while _FORK_ON and len(pool) < n - 1:
            pool.append((torch.cuda.Stream(device=device), torch.cuda.Event(), torch.cuda.Event()))
        self.side = pool[:n - 1] if _FORK_ON else []

    def begin(self):
        for s, ev_fork, _ in self.side:
            ev_fork.record(self.main)
            s.wait_event(ev_fork)

    def on(self, i):
        """Context: item 0 stays on the main stream, item i > 0 runs on side stream i - 1."""
        if i == 0 or not self.side:
            return contextlib.nullcontext()
        return torch.cuda.stream(self.side[i - 1][0])

    def join(self):
        for s, _, ev_join in self.side:
            ev_join.record(s)
            self.main.wait_event(ev_join)



def ballquery(coords, batch_idxs, batch_offsets, radius):
    """Returns (idx[nActive] i32, start_len[n,2] i32); one host read of nActive."""
    require_cuda(coords, batch_idxs, batch_offsets)
    require(coords.dtype == torch.float32 and coords.is_contiguous() and coords.dim() == 2 and coords.size(1) == 3,
            "coords must be contiguous float32 [n,3]")
    require(batch_idxs.dtype == torch.uint8 and batch_idxs.is_contiguous(), "batch_idxs must be contiguous uint8")
    require(batch_offsets.dtype == I32 and batch_offsets.is_contiguous(), "batch_offsets must be contiguous int32")
    n = coords.size(0)
    dev = coords.device
    start_len = torch.empty((n, 2), dtype=I32, device=dev)
    d_count = _dev_i32(1, dev)
    ws = workspace(lib().b2s_ballquery_ws_bytes(n), dev)
    nb = batch_offsets.numel() - 1
    check(lib().b2s_ballquery_count(ptr(coords), ptr(batch_idxs), ptr(batch_offsets), n, nb, float(radius),
                                    ptr(start_len), ptr(d_count), ptr(ws), ws.numel(), stream()), "ballquery_count")
    n_active = int(d_count.item()) if n > 0 else 0
    run_deferred_checks()
    idx = _dev_i32(n_active, dev)
    if n_active > 0:
        check(lib().b2s_ballquery_fill(ptr(coords), ptr(batch_idxs), ptr(batch_offsets), n, nb, float(radius),
                                       ptr(start_len), ptr(idx), ptr(ws), ws.numel(), stream()), "ballquery_fill")
    return idx, start_len


def ballquery_many(coord_sets, batch_idxs, batch_offsets, radius, between=None):
    """Several ball queries over the same points' batch layout (PointGroup: raw and shifted coordinates,
    pointgroup.py:43,58) with ONE host read for all pair counts instead of one per query: all count passes are
    enqueued first (each keeps its own cell grid), the counts are read together, then all fill passes run.
    Returns [(idx, start_len), ...], identical to [ballquery(c, ...) for c in coord_sets].
    `between` (optional callable) runs after the count passes are enqueued and before their results are read."""
    require_cuda(batch_idxs, batch_offsets, *coord_sets)
    n = batch_idxs.numel()
    dev = batch_idxs.device
    nb = batch_offsets.numel() - 1
    if n == 0:
        return [ballquery(c, batch_idxs, batch_offsets, radius) for c in coord_sets]
    ws_bytes = lib().b2s_ballquery_ws_bytes(n)
    state = []
    fork = _Fork(dev, len(coord_sets))
    fork.begin()
    for qi, coords in enumerate(coord_sets):
        require(coords.dtype == torch.float32 and coords.is_contiguous() and tuple(coords.shape) == (n, 3),
                "coords must be contiguous float32 [n,3]")
        start_len = torch.empty((n, 2), dtype=I32, device=dev)
        d_count = _dev_i32(1, dev)
        with fork.on(qi):
            ws = workspace(ws_bytes, dev, slot=1 + qi)  # private: the cell grid must survive until the fill
            check(lib().b2s_ballquery_count(ptr(coords), ptr(batch_idxs), ptr(batch_offsets), n, nb, float(radius),
                                            ptr(start_len), ptr(d_count), ptr(ws), ws.numel(), stream()),
                  "ballquery_count")
        state.append((coords, start_len, d_count, ws))
    fork.join()
    d_counts = torch.cat([st[2] for st in state])
    if between is not None:
        between()  # independent host work of the caller, enqueued while the count passes run
    counts = d_counts.tolist()
    run_deferred_checks()
    out = []
    fork.begin()
    for qi, ((coords, start_len, _, ws), n_active) in enumerate(zip(state, counts)):
        idx = _dev_i32(n_active, dev)
        if n_active > 0:
            with fork.on(qi):
                check(lib().b2s_ballquery_fill(ptr(coords), ptr(batch_idxs), ptr(batch_offsets), n, nb, float(radius),
                                               ptr(start_len), ptr(idx), ptr(ws), ws.numel(), stream()),
                      "ballquery_fill")
        out.append((idx, start_len))
    fork.join()
    return out


# ------------------------------------------------------------------------------------------
# C2 / C3 / C4 clustering
# ------------------------------------------------------------------------------------------
def cluster_label(nbr_idx, start_len, labels=None):
    require_cuda(nbr_idx, start_len, labels)
    require(nbr_idx.dtype == I32 and start_len.dtype == I32 and nbr_idx.is_contiguous() and start_len.is_contiguous(),
            "ball query outputs must be contiguous int32")
    if labels is not None:
        require(labels.dtype == torch.int16 and labels.is_contiguous(), "labels must be contiguous int16")
    n = start_len.size(0)
    comp = _dev_i32(n, start_len.device)
    ws = workspace(lib().b2s_cluster_ws_bytes(n), start_len.device)
    check(lib().b2s_cluster_label(ptr(nbr_idx), ptr(start_len), ptr(labels), n, ptr(comp), ptr(ws), ws.numel(),
                                  stream()), "cluster_label")
    return comp


def cluster_extract(nbr_idx, start_len, labels, comp, mode, thr_i=0, thr_f=0.0, point_num_avg=None, group=0):
    """select + order: returns (cluster_idxs[S,2] i32, cluster_offsets[nC+1] i32). One host read."""
    n = start_len.size(0)
    dev = start_len.device
    offsets = _dev_i32(n + 1, dev)
    seeds = _dev_i32(max(n, 1), dev)
    d_count = _dev_i32(2, dev)
    ws = workspace(lib().b2s_cluster_ws_bytes(n), dev)
    check(lib().b2s_cluster_select(ptr(comp), ptr(labels), n, mode, int(thr_i), float(thr_f), ptr(point_num_avg),
                                   group, ptr(offsets), ptr(seeds), ptr(d_count), ptr(ws), ws.numel(), stream()),
          "cluster_select")
    n_cluster, total = d_count.tolist()
    run_deferred_checks()
    cluster_idxs = torch.empty((total, 2), dtype=I32, device=dev)
    offsets = offsets[:n_cluster + 1]
    if n_cluster > 0:
        check(lib().b2s_cluster_order(ptr(nbr_idx), ptr(start_len), ptr(labels), ptr(comp), n, nbr_idx.numel(),
                                      ptr(offsets),
                                      ptr(seeds), n_cluster, ptr(cluster_idxs), ptr(ws), ws.numel(), stream()),
              "cluster_order")
    return cluster_idxs, offsets


def pg_cluster_many(labels, queries, threshold):
    """pg_bfs_cluster (mode 0) on several ball-query results over the same labelled points with ONE host read for all
    (nCluster, sumNPoint) pairs: label + select of every query are enqueued first (private scratch each), the counts
    are read together, then the BFS orders run.  Returns [(cluster_idxs, cluster_offsets), ...]."""
    require_cuda(labels)
    n = labels.numel()
    dev = labels.device
    ws_bytes = lib().b2s_cluster_ws_bytes(n)
    state = []
    fork = _Fork(dev, len(queries))
    fork.begin()
    for qi, (nbr_idx, start_len) in enumerate(queries):
        comp = _dev_i32(n, dev)
        offsets = _dev_i32(n + 1, dev)
        seeds = _dev_i32(max(n, 1), dev)
        d_count = _dev_i32(2, dev)
        with fork.on(qi):
            ws = workspace(ws_bytes, dev, slot=1 + qi)
            check(lib().b2s_cluster_label(ptr(nbr_idx), ptr(start_len), ptr(labels), n, ptr(comp), ptr(ws), ws.numel(),
                                          stream()), "cluster_label")
            check(lib().b2s_cluster_select(ptr(comp), ptr(labels), n, 0, int(threshold), 0.0, None, 0, ptr(offsets),
                                           ptr(seeds), ptr(d_count), ptr(ws), ws.numel(), stream()), "cluster_select")
        state.append((nbr_idx, start_len, comp, offsets, seeds, d_count, ws))
    fork.join()
    counts = torch.stack([st[5] for st in state]).tolist()
    run_deferred_checks()
    out = []
    fork.begin()
    for qi, ((nbr_idx, start_len, comp, offsets, seeds, _, ws), (n_cluster, total)) in enumerate(zip(state, counts)):
        cluster_idxs = torch.empty((total, 2), dtype=I32, device=dev)
        offsets = offsets[:n_cluster + 1]
        if n_cluster > 0:
            with fork.on(qi):
                check(lib().b2s_cluster_order(ptr(nbr_idx), ptr(start_len), ptr(labels), ptr(comp), n, nbr_idx.numel(),
                                              ptr(offsets), ptr(seeds), n_cluster, ptr(cluster_idxs), ptr(ws),
                                              ws.numel(), stream()), "cluster_order")
        out.append((cluster_idxs, offsets))
    fork.join()
    return out


def cluster_centers(cluster_idxs, cluster_offsets, coords, labels, batch_idxs):
    nc = cluster_offsets.numel() - 1
    centers = torch.empty((nc, 5), dtype=torch.float32, device=coords.device)
    check(lib().b2s_cluster_centers(ptr(cluster_idxs), ptr(cluster_offsets), nc, ptr(coords), ptr(labels),
                                    ptr(batch_idxs), ptr(centers), stream()), "cluster_centers")
    return centers


def ha_set_aggregate(frag_idxs, frag_offsets, frag_centers, prim_idxs, prim_offsets, prim_centers, radius_avg):
    """HAIS set aggregation; returns (primary_idxs_post[S',2], primary_offsets_post[nP+1])."""
    dev = prim_idxs.device
    n_frag = frag_offsets.numel() - 1
    n_prim = prim_offsets.numel() - 1
    assign = torch.full((max(n_frag, 1),), -1, dtype=I32, device=dev)
    check(lib().b2s_ha_assign(ptr(frag_centers), n_frag, ptr(prim_centers), ptr(prim_offsets), n_prim,
                              ptr(radius_avg), ptr(assign), stream()), "ha_assign")
    out_idxs = torch.empty((frag_idxs.size(0) + prim_idxs.size(0), 2), dtype=I32, device=dev)
    out_offsets = torch.zeros(n_prim + 1, dtype=I32, device=dev)
    ws = workspace(lib().b2s_ha_concat_ws_bytes(n_frag, n_prim), dev)
    check(lib().b2s_ha_concat(ptr(frag_idxs), ptr(frag_offsets), n_frag, ptr(prim_idxs), ptr(prim_offsets), n_prim,
                              ptr(assign), ptr(out_idxs), ptr(out_offsets), ptr(ws), ws.numel(), stream()),
          "ha_concat")
    return out_idxs, out_offsets, assign[:n_frag]


# ------------------------------------------------------------------------------------------
# S1-S3, I1-I2
# ------------------------------------------------------------------------------------------
def _seg(fn_name, inp, offsets, out, *extra):
    n_seg = offsets.numel() - 1
    c = inp.size(1)
    fn = getattr(lib(), fn_name)
    check(fn(ptr(inp), ptr(offsets), ptr(out), *[ptr(e) for e in extra], n_seg, c, stream()), fn_name)


def sec_reduce(kind, inp, offsets, out):
    _f32c(inp, "inp")
    require_cuda(offsets, out)
    require(offsets.dtype == I32 and offsets.is_contiguous(), "offsets must be contiguous int32")
    _seg({"mean": "b2s_sec_mean", "min": "b2s_sec_min", "max": "b2s_sec_max",
          "avg": "b2s_global_avg_pool_fp"}[kind], inp, offsets, out)
    return out


def roipool_fp(feats, offsets, out, maxidx):
    _f32c(feats, "feats")
    require(offsets.dtype == I32 and offsets.is_contiguous(), "offsets must be contiguous int32")
    n_seg, c = offsets.numel() - 1, feats.size(1)
    ws = workspace(lib().b2s_roipool_ws_bytes(n_seg, c), feats.device)
    check(lib().b2s_roipool_fp_ws(ptr(feats), ptr(offsets), ptr(out), ptr(maxidx), n_seg, c, feats.size(0), ptr(ws),
                                  ws.numel(), stream()), "roipool_fp")


def roipool_bp(d_feats, offsets, maxidx, d_out):
    n_seg, c = d_out.shape
    check(lib().b2s_roipool_bp(ptr(d_feats), ptr(offsets), ptr(maxidx), ptr(d_out), n_seg, c, stream()), "roipool_bp")


def global_avg_pool_bp(d_feats, offsets, d_out):
    n_seg, c = d_out.shape
    check(lib().b2s_global_avg_pool_bp(ptr(d_feats), ptr(offsets), ptr(d_out), n_seg, c, stream()),
          "global_avg_pool_bp")


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, mask_scores=None):
    require_cuda(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, mask_scores)
    require(proposals_idx.dtype == I32 and proposals_offset.dtype == I32 and instance_pointnum.dtype == I32,
            "proposals_idx / proposals_offset / instance_pointnum must be int32")
    require(instance_labels.dtype == torch.int16, "instance labels must be int16")
    n_inst = instance_pointnum.numel()
    n_prop = proposals_offset.numel() - 1
    check(lib().b2s_get_iou(ptr(proposals_idx), ptr(proposals_offset), ptr(instance_labels), ptr(instance_pointnum),
                            ptr(mask_scores), ptr(proposals_iou), n_inst, n_prop, stream()), "get_iou")
    return proposals_iou


def get_mask_label(proposals_idx, proposals_offset, instance_labels, instance_cls, proposals_iou, n_inst, n_prop,
                   ignored_label, iou_thr, mask_label, mask_label_mask):
    require_cuda(proposals_idx, proposals_offset, instance_labels, instance_cls, proposals_iou)
    require(instance_cls.dtype == torch.int16 and instance_labels.dtype == torch.int16, "labels must be int16")
    check(lib().b2s_get_mask_label(ptr(proposals_idx), ptr(proposals_offset), ptr(instance_labels),
                                   ptr(instance_cls), ptr(proposals_iou), n_inst, n_prop, int(ignored_label),
                                   float(iou_thr), ptr(mask_label), ptr(mask_label_mask), stream()), "get_mask_label")


# ------------------------------------------------------------------------------------------
# semantic cross-entropy of the train step (general_model.py:36-40)
# ------------------------------------------------------------------------------------------
class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, labels, ignore_index):
        _f32c(scores, "scores")
        require_cuda(labels)
        require(labels.dtype in (torch.int16, torch.int64) and labels.is_contiguous() and labels.numel() == scores.size(0),
                "labels must be contiguous int16 / int64 [n]")
        n, c = scores.shape
        out = torch.empty(2, dtype=torch.float32, device=scores.device)  # loss, n_valid
        ws = workspace(lib().b2s_cross_entropy_ws_bytes(n), scores.device)
        check(lib().b2s_cross_entropy_forward(ptr(scores), ptr(labels), labels.element_size(), n, c, int(ignore_index),
                                              ptr(out), out.data_ptr() + 4, ptr(_bn_counter(scores.device)), ptr(ws),
                                              ws.numel(), stream()), "cross_entropy_forward")
        ctx.save_for_backward(scores, labels, out)
        ctx.ignore_index = int(ignore_index)
        return out[0]

    @staticmethod
    def backward(ctx, gout):
        scores, labels, out = ctx.saved_tensors
        n, c = scores.shape
        g = torch.empty_like(scores)
        gout = gout.contiguous().to(torch.float32).reshape(1)
        check(lib().b2s_cross_entropy_backward(ptr(scores), ptr(labels), labels.element_size(), n, c, ctx.ignore_index,
                                               out.data_ptr() + 4, ptr(gout), ptr(g), stream()), "cross_entropy_backward")
        return g, None, None


def cross_entropy(scores, labels, ignore_index=-1):
    """F.cross_entropy(scores, labels.long(), ignore_index=ignore_index) (mean over the labelled rows) for float32
    [n, c] CUDA scores and int16 / int64 labels, fused forward and backward kernels."""
    if scores.size(0) == 0:
        return torch.nn.functional.cross_entropy(scores, labels.long(), ignore_index=ignore_index)
    return _CrossEntropy.apply(scores.contiguous(), labels.contiguous(), ignore_index)


# ------------------------------------------------------------------------------------------
# clusters_voxelization up to `.int()` (general_model.py:152-182)
# ------------------------------------------------------------------------------------------
def clusters_voxelize(clusters_idx, clusters_offset, coords, scale, spatial_shape, rand):
    """Integer voxel coordinates [sumNPoint, 4] = (cluster id, x, y, z) of every proposal point.

    clusters_idx [S, 2] int32/int64, clusters_offset [nC+1] int32, coords [N, 3] float32, rand [2, 3] float32 (the
    two torch.rand(3) draws of the reference).  Bit-identical to the reference's torch expression sequence.
    """
    require_cuda(clusters_idx, clusters_offset, coords, rand)
    require(clusters_idx.dim() == 2 and clusters_idx.size(1) == 2 and clusters_idx.is_contiguous()
            and clusters_idx.dtype in (torch.int32, torch.int64), "clusters_idx must be contiguous int32/int64 [S, 2]")
    require(clusters_offset.dtype == I32 and clusters_offset.is_contiguous(), "clusters_offset must be int32")
    _f32c(coords, "coords")
    rand = rand.to(torch.float32).contiguous()
    s_total, n_cluster = clusters_idx.size(0), clusters_offset.numel() - 1
    out = torch.empty((s_total, 4), dtype=I32, device=coords.device)
    params = torch.empty((max(n_cluster, 1), 8), dtype=torch.float32, device=coords.device)
    check(lib().b2s_clusters_voxelize(ptr(clusters_idx), int(clusters_idx.dtype == torch.int64), ptr(clusters_offset),
                                      s_total, n_cluster, ptr(coords), float(scale), int(spatial_shape), ptr(rand),
                                      ptr(out), ptr(params), stream()), "clusters_voxelize")
    return out


__all__ = [n for n in dir() if not n.startswith("_")]
