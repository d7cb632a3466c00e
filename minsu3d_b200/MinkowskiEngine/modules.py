"""nn.Module surface: MinkowskiConvolution(+Transpose), MinkowskiBatchNorm, MinkowskiReLU, cat.

Forward/backward algebra:
    conv:    out[o] = sum_k in[nbr[o,k]] @ kernel[k]
             gin[i] = sum_k gout[nbr_T[i,k]] @ kernel[k]^T ;  gkernel[k] = in[I_k]^T @ gout[O_k]
    convT:   forward strided map with in/out swapped and the same kernel index.
The products run in libb2s (include/b2s.h: b2s_conv_table / b2s_conv_pairs / b2s_conv_wgrad).
"""
import math

import torch
import torch.nn as nn

from .. import ops
from .sparse_tensor import SparseTensor


# ------------------------------------------------------------------------------------------
# autograd functions (raw feature matrices + cached kernel maps)
# ------------------------------------------------------------------------------------------
class _ConvSame(torch.autograd.Function):
    """Stride-1 odd-size convolution on one coordinate map: symmetric kernel map."""

    @staticmethod
    def forward(ctx, feats, kernel, kmap, packed=None):
        feats = feats.contiguous()
        K, cin, cout = kernel.shape
        nbr, tile_mask, out_rows = kmap.table_for(cin, cout)
        out = ops.conv_table(feats, kernel, nbr, kmap.n_out, K, cin, cout, tile_mask=tile_mask, out_rows=out_rows,
                             packed=packed)
        ctx.save_for_backward(feats, kernel)
        ctx.kmap = kmap
        ctx.packed = packed
        return out

    @staticmethod
    def backward(ctx, gout):
        feats, kernel = ctx.saved_tensors
        kmap = ctx.kmap
        K, cin, cout = kernel.shape
        gout = gout.contiguous()
        gin = gk = None
        if ctx.needs_input_grad[0]:
            # nbr[i, K-1-k] = o  <=>  nbr[o, k] = i: reuse the table with reversed, transposed weights
            nbr, tile_mask, out_rows = kmap.table_for(cout, cin)
            gin = ops.conv_table(gout, kernel, nbr, kmap.n_in, K, cout, cin, w_transposed=True, k_reversed=True,
                                 tile_mask=tile_mask, out_rows=out_rows, packed=ctx.packed)
        if ctx.needs_input_grad[1]:
            pin, pout, koff, maxp = kmap.pairs()
            gk = ops.conv_wgrad(feats, gout, pin, pout, koff, K, cin, cout, maxp)
        return gin, gk, None, None


class _ConvDown(torch.autograd.Function):
    """kernel_size == stride (2) convolution: fine -> coarse, table nbr[coarse, 8] of fine rows."""

    @staticmethod
    def forward(ctx, feats, kernel, kmap, packed=None):
        feats = feats.contiguous()
        K, cin, cout = kernel.shape
        out = ops.conv_table(feats, kernel, kmap.nbr, kmap.n_out, K, cin, cout, tile_mask=kmap.tile_mask, packed=packed)
        ctx.save_for_backward(feats, kernel)
        ctx.kmap = kmap
        ctx.packed = packed
        return out

    @staticmethod
    def backward(ctx, gout):
        feats, kernel = ctx.saved_tensors
        kmap = ctx.kmap
        K, cin, cout = kernel.shape
        gout = gout.contiguous()
        pin, pout, koff, maxp = kmap.pairs()
        gin = gk = None
        if ctx.needs_input_grad[0]:
            # every fine row has exactly one (coarse row, offset): plain store, no accumulation
            gin = ops.conv_pairs(gout, kernel, pout, pin, koff, kmap.n_in, K, cout, cin, maxp, w_transposed=True,
                                 packed=ctx.packed)
        if ctx.needs_input_grad[1]:
            gk = ops.conv_wgrad(feats, gout, pin, pout, koff, K, cin, cout, maxp)
        return gin, gk, None, None


class _ConvUp(torch.autograd.Function):
    """Transposed kernel_size == stride (2) convolution: coarse -> fine on the existing fine map.

    kmap is the FORWARD strided map (fine -> coarse); it is used with in/out swapped.
    """

    @staticmethod
    def forward(ctx, feats, kernel, kmap, packed=None):
        feats = feats.contiguous()
        K, cin, cout = kernel.shape
        pin, pout, koff, maxp = kmap.pairs()
        out = ops.conv_pairs(feats, kernel, pout, pin, koff, kmap.n_in, K, cin, cout, maxp, packed=packed)
        ctx.save_for_backward(feats, kernel)
        ctx.kmap = kmap
        ctx.packed = packed
        return out

    @staticmethod
    def backward(ctx, gout):
        feats, kernel = ctx.saved_tensors
        kmap = ctx.kmap
        K, cin, cout = kernel.shape
        gout = gout.contiguous()
        gin = gk = None
        if ctx.needs_input_grad[0]:
            gin = ops.conv_table(gout, kernel, kmap.nbr, kmap.n_out, K, cout, cin, w_transposed=True,
                                 tile_mask=kmap.tile_mask, packed=ctx.packed)
        if ctx.needs_input_grad[1]:
            pin, pout, koff, maxp = kmap.pairs()
            gk = ops.conv_wgrad(feats, gout, pout, pin, koff, K, cin, cout, maxp)
        return gin, gk, None, None


_IDENT = {}


def _identity_pairs(n, device):
    """(arange(n) i32, k_offsets [0, n]) for the dense 1x1 case, cached per size."""
    key = (n, device.index)
    v = _IDENT.get(key)
    if v is None:
        if len(_IDENT) > 64:
            _IDENT.clear()
        v = _IDENT[key] = (torch.arange(n, dtype=torch.int32, device=device),
                           torch.tensor([0, n], dtype=torch.int32, device=device))
    return v


class _Conv1x1(torch.autograd.Function):
    """kernel_size 1, stride 1: dense [M,Cin] @ [Cin,Cout] through the same implicit-GEMM kernel."""

    @staticmethod
    def forward(ctx, feats, kernel, packed=None):
        feats = feats.contiguous()
        cin, cout = kernel.shape
        out = ops.conv_table(feats, kernel, None, feats.size(0), 1, cin, cout, packed=packed)
        ctx.save_for_backward(feats, kernel)
        ctx.packed = packed
        return out

    @staticmethod
    def backward(ctx, gout):
        feats, kernel = ctx.saved_tensors
        cin, cout = kernel.shape
        gout = gout.contiguous()
        gin = gk = None
        if ctx.needs_input_grad[0]:
            gin = ops.conv_table(gout, kernel, None, feats.size(0), 1, cout, cin, w_transposed=True, packed=ctx.packed)
        if ctx.needs_input_grad[1]:
            n = feats.size(0)
            ident, koff = _identity_pairs(n, feats.device)
            gk = ops.conv_wgrad(feats, gout, ident, ident, koff, 1, cin, cout, n).view(cin, cout)
        return gin, gk, None


# ------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------
def _scalar(v, name):
    if isinstance(v, (list, tuple)):
        if len(set(v)) != 1:
            raise NotImplementedError("anisotropic %s is not supported" % name)
        v = v[0]
    return int(v)


class _ConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, is_transpose=False, expand_coordinates=False, dimension=None, **_unused):
        super().__init__()
        if dimension != 3:
            raise NotImplementedError("only dimension=3 is supported (every minsu3d call site uses 3)")
        if kernel_generator is not None or expand_coordinates:
            raise NotImplementedError("custom kernel generators / expand_coordinates are not supported")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _scalar(kernel_size, "kernel_size")
        self.stride = _scalar(stride, "stride")
        self.dilation = _scalar(dilation, "dilation")
        self.is_transpose = is_transpose
        self.dimension = dimension
        if self.dilation != 1:
            raise NotImplementedError("dilation != 1 is not used by minsu3d and not supported")
        self.kernel_volume = self.kernel_size ** 3
        self.use_mm = self.kernel_volume == 1 and self.stride == 1
        if self.use_mm:
            shape = (in_channels, out_channels)
        else:
            shape = (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape, dtype=torch.float32))
        self.bias = nn.Parameter(torch.empty(1, out_channels, dtype=torch.float32)) if bias else None
        self._packed = ops.PackedWeights()  # tensor-core operand images of `kernel`, re-packed when it changes
        self.reset_parameters(is_transpose)

    def packed(self):
        """conv_pack(kernel), cached until the parameter is modified (None: shape not on the tcgen05 path / CPU)."""
        if ops.get_conv_algo() == ops.ALGO_SIMT:
            return None
        return self._packed.get(self.kernel)

    def reset_parameters(self, is_transpose=False):
        with torch.no_grad():
            n = (self.out_channels if is_transpose else self.in_channels) * self.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def _finish(self, x, feats, key):
        if self.bias is not None:
            feats = feats + self.bias
        return SparseTensor(feats, coordinate_map_key=key, coordinate_manager=x.coordinate_manager)

    def extra_repr(self):
        return "in=%d, out=%d, kernel_size=%d, stride=%d" % (self.in_channels, self.out_channels,
                                                              self.kernel_size, self.stride)


class MinkowskiConvolution(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator,
                         False, expand_coordinates, dimension)

    def forward(self, x):
        mgr, key = x.coordinate_manager, x.coordinate_map_key
        if self.use_mm:
            return self._finish(x, _Conv1x1.apply(x.F, self.kernel, self.packed()), key)
        if self.stride == 1:
            if self.kernel_size % 2 != 1:
                raise NotImplementedError("stride-1 convolutions need an odd kernel size")
            kmap = mgr.kernel_map(key, key, self.kernel_size)
            return self._finish(x, _ConvSame.apply(x.F, self.kernel, kmap, self.packed()), key)
        if self.stride != self.kernel_size:
            raise NotImplementedError("strided convolution is supported for kernel_size == stride (2/2)")
        out_key = mgr.stride_key(key, self.stride)
        kmap = mgr.kernel_map(key, out_key, self.kernel_size)
        return self._finish(x, _ConvDown.apply(x.F, self.kernel, kmap, self.packed()), out_key)


class MinkowskiConvolutionTranspose(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator,
                         True, expand_coordinates, dimension)

    def forward(self, x):
        mgr, key = x.coordinate_manager, x.coordinate_map_key
        if self.use_mm:
            return self._finish(x, _Conv1x1.apply(x.F, self.kernel, self.packed()), key)
        if self.stride != self.kernel_size or key.stride % self.stride != 0:
            raise NotImplementedError("transposed convolution is supported for kernel_size == stride (2/2)")
        fine_key = mgr.existing_key(key.stride // self.stride)  # the encoder's map (appendix A.6)
        kmap = mgr.kernel_map(fine_key, key, self.kernel_size)
        return self._finish(x, _ConvUp.apply(x.F, self.kernel, kmap, self.packed()), fine_key)


# ---- normalisation / activation -------------------------------------------------------------
class _BNReLU(torch.autograd.Function):
    """y = relu?(batch_norm(x)) with libb2s kernels: stats pass + fused apply, fused backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, relu):
        x = x.contiguous()
        if training:
            y, mean, rstd = ops.bn_forward(x, eps, momentum if running_mean is not None else 0.0, running_mean,
                                           running_var, weight, bias, relu)
        else:
            mean, rstd = running_mean, torch.rsqrt(running_var + eps)
            y = ops.bn_apply(x, mean, rstd, weight, bias, relu)
        ctx.save_for_backward(x, y if relu else x, mean, rstd, weight)
        ctx.relu = relu
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, weight = ctx.saved_tensors
        dx, dgamma, dbeta = ops.bn_backward(x, y, dy.contiguous(), mean, rstd, weight, ctx.relu, ctx.training)
        return dx, dgamma, dbeta, None, None, None, None, None, None


class MinkowskiBatchNorm(nn.Module):
    """nn.BatchNorm1d on the feature matrix; child module is named `bn` like MinkowskiEngine's."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def _features(self, feats, relu):
        bn = self.bn
        c = feats.size(1)
        use_batch = bn.training or not bn.track_running_stats
        if not bn.affine or c % 4 != 0 or feats.size(0) == 0 or (bn.momentum is None and use_batch):
            y = bn(feats)
            return torch.relu(y) if relu else y
        if use_batch and bn.track_running_stats:
            bn.num_batches_tracked += 1
        return _BNReLU.apply(feats, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch,
                             bn.momentum, bn.eps, relu)

    def forward(self, x):
        out = x._like(None)
        out._pending_bn = (self, x.F)  # materialised on first access; fused with a following ReLU
        return _LazyBN.wrap(out)

    def __repr__(self):
        return "MinkowskiBatchNorm(%d, eps=%g, momentum=%s)" % (self.bn.num_features, self.bn.eps, self.bn.momentum)


class _LazyBN:
    """Defers BN so that BN -> ReLU (every use in minsu3d: common.py:13-14,35-39,67-68,75-76;
    backbone.py:16-17) runs as one fused apply kernel instead of two passes."""

    @staticmethod
    def wrap(st):
        st.__class__ = _PendingSparseTensor
        return st


class _PendingSparseTensor(SparseTensor):
    def _materialise(self, relu):
        mod, feats = self._pending_bn
        self._pending_bn = None
        self._F = mod._features(feats, relu)
        self.__class__ = SparseTensor
        return self

    @property
    def F(self):
        return self._materialise(False)._F

    @property
    def features(self):
        return self._materialise(False)._F

    @features.setter
    def features(self, value):
        self._pending_bn = None
        self._F = value
        self.__class__ = SparseTensor

    @property
    def device(self):
        return self._pending_bn[1].device

    @property
    def dtype(self):
        return self._pending_bn[1].dtype

    @property
    def shape(self):
        return self._pending_bn[1].shape

    def size(self, *a):
        return self._pending_bn[1].size(*a)

    def __len__(self):
        return self._pending_bn[1].size(0)

    def __add__(self, other):
        return self._materialise(False).__add__(other)

    def __iadd__(self, other):
        return self._materialise(False).__iadd__(other)

    def __sub__(self, other):
        return self._materialise(False).__sub__(other)

    def __mul__(self, other):
        return self._materialise(False).__mul__(other)


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        if isinstance(x, _PendingSparseTensor):
            return x._materialise(True)
        return x._like(torch.relu(x.F))

    def __repr__(self):
        return "MinkowskiReLU()"


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x):
        return x._like(self.linear(x.F))


class MinkowskiGlobalAvgPooling(nn.Module):
    """Per-batch-index mean of the features (unused by minsu3d; thin alias over the S3 kernel)."""

    def forward(self, x):
        from ..common_ops.functions.softgroup_ops import global_avg_pool
        idx = x.C[:, 0].long()
        order = torch.argsort(idx, stable=True)
        offsets = torch.cumsum(torch.bincount(idx + 1), dim=0).int()
        return global_avg_pool(x.F[order].contiguous(), offsets)


def cat(*sparse_tensors):
    """Channel concatenation of tensors on the same coordinate map (common.py:93)."""
    if len(sparse_tensors) == 1 and isinstance(sparse_tensors[0], (list, tuple)):
        sparse_tensors = tuple(sparse_tensors[0])
    first = sparse_tensors[0]
    for s in sparse_tensors[1:]:
        first._check_same_map(s)
    return first._like(torch.cat([s.F for s in sparse_tensors], dim=1))


# ------------------------------------------------------------------------------------------
# host-side fusion: one residual block = one autograd node + one library call (csrc/fused.cu)
# ------------------------------------------------------------------------------------------
class _ResBlockFn(torch.autograd.Function):
    """out = conv3(relu(bn2(conv3(relu(bn1(x)))))) + shortcut(x) in training mode; bit-identical to the
    module-by-module path (same kernels in the same order), but 1 node / 1 call instead of 5-6 each."""

    @staticmethod
    def forward(ctx, x, g1, b1, w1, g2, b2, w2, wds, bn1, bn2, kmap, wp1=None, wp2=None, wpds=None):
        x = x.contiguous()
        n, cin = x.shape
        cout = w1.size(2)
        K = w1.size(0)
        dev = x.device
        # one buffer for everything the backward needs: y1 [n,cin] | z1 [n,cout] | y2 [n,cout] | stats1 | stats2
        saved = torch.empty(n * (cin + 2 * cout) + 2 * cin + 2 * cout, dtype=torch.float32, device=dev)
        out = torch.empty((n, cout), dtype=torch.float32, device=dev)
        tmp = torch.empty((n, cout), dtype=torch.float32, device=dev) if wds is not None else None
        base = saved.data_ptr()
        p_y1, p_z1 = base, base + 4 * n * cin
        p_y2 = p_z1 + 4 * n * cout
        p_s1 = p_y2 + 4 * n * cout
        p_s2 = p_s1 + 8 * cin
        nbr_s, mask_s, perm = kmap.sorted_tables(cin, cout)
        algo = ops.get_conv_algo()
        ws = ops.workspace(ops.resblock_ws_bytes(K, cin, cout), dev)
        ops.check(ops.lib().b2s_resblock_forward(
            x.data_ptr(), n, cin, cout,
            g1.data_ptr(), b1.data_ptr(), ops.ptr(bn1.running_mean), ops.ptr(bn1.running_var), w1.data_ptr(),
            g2.data_ptr(), b2.data_ptr(), ops.ptr(bn2.running_mean), ops.ptr(bn2.running_var), w2.data_ptr(),
            ops.ptr(wds), ops.ptr(wp1), ops.ptr(wp2), ops.ptr(wpds),
            bn1.eps, bn1.momentum if bn1.running_mean is not None else 0.0,
            bn2.eps, bn2.momentum if bn2.running_mean is not None else 0.0,
            kmap.nbr.data_ptr(), ops.ptr(kmap.tile_mask), ops.ptr(nbr_s), ops.ptr(mask_s), ops.ptr(perm), K,
            p_y1, p_s1, p_z1, p_y2, p_s2, out.data_ptr(), ops.ptr(tmp),
            ops.bn_counter(dev).data_ptr(), algo, ws.data_ptr(), ws.numel(), ops.stream()), "resblock_forward")
        ctx.save_for_backward(x, saved, g1, g2, w1, w2, wds)
        ctx.kmap = kmap
        ctx.algo = algo
        ctx.packed = (wp1, wp2, wpds)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, saved, g1, g2, w1, w2, wds = ctx.saved_tensors
        kmap = ctx.kmap
        gout = gout.contiguous()
        n, cin = x.shape
        K, _, cout = w1.shape
        dev = x.device
        base = saved.data_ptr()
        p_y1, p_z1 = base, base + 4 * n * cin
        p_y2 = p_z1 + 4 * n * cout
        p_s1 = p_y2 + 4 * n * cout
        p_s2 = p_s1 + 8 * cin
        gx = torch.empty_like(x)
        gw1 = torch.empty_like(w1)
        gw2 = torch.empty_like(w2)
        gwds = torch.empty_like(wds) if wds is not None else None
        dgb1 = torch.empty((2, cin), dtype=torch.float32, device=dev)
        dgb2 = torch.empty((2, cout), dtype=torch.float32, device=dev)
        wp1, wp2, wpds = ctx.packed
        scratch = torch.empty(n * (max(cin, cout) + cout + cin), dtype=torch.float32, device=dev)
        p_a = scratch.data_ptr()  # [n, max(cin, cout)]: conv2 data gradient, later the 1x1 shortcut's data gradient
        p_b = p_a + 4 * n * max(cin, cout)
        p_c = p_b + 4 * n * cout
        pin, pout, koff, maxp = kmap.pairs()
        ident = ident_koff = None
        if wds is not None:
            ident, ident_koff = _identity_pairs(n, dev)
        nbr_s, mask_s, perm = kmap.sorted_tables(cin, cout)
        ws = ops.workspace(ops.resblock_ws_bytes(K, cin, cout), dev)
        ops.check(ops.lib().b2s_resblock_backward(
            gout.data_ptr(), x.data_ptr(), p_y1, p_z1, p_y2, p_s1, p_s2, g1.data_ptr(), g2.data_ptr(),
            w1.data_ptr(), w2.data_ptr(), ops.ptr(wds), ops.ptr(wp1), ops.ptr(wp2), ops.ptr(wpds), n, cin, cout,
            kmap.nbr.data_ptr(), ops.ptr(kmap.tile_mask), ops.ptr(nbr_s), ops.ptr(mask_s), ops.ptr(perm), K,
            pin.data_ptr(), pout.data_ptr(), koff.data_ptr(), int(maxp), ops.ptr(ident), ops.ptr(ident_koff),
            gx.data_ptr(), gw1.data_ptr(), gw2.data_ptr(), ops.ptr(gwds), dgb1.data_ptr(), dgb2.data_ptr(),
            p_a, p_b, p_c, ops.bn_counter(dev).data_ptr(), ctx.algo, ws.data_ptr(), ws.numel(), ops.stream()),
            "resblock_backward")
        return gx, dgb1[0], dgb1[1], gw1, dgb2[0], dgb2[1], gw2, gwds, None, None, None, None, None, None


def residual_block_fusable(x, bn1, conv1, bn2, conv2, downsample):
    """Training-mode pre-activation residual block on one coordinate map with 3^3 stride-1 convolutions."""
    b1, b2 = bn1.bn, bn2.bn
    return (b1.training and b2.training and b1.affine and b2.affine and b1.track_running_stats
            and b2.track_running_stats and b1.momentum is not None and b2.momentum is not None
            and conv1.kernel_size == 3 and conv2.kernel_size == 3 and conv1.stride == 1 and conv2.stride == 1
            and conv1.bias is None and conv2.bias is None and not conv1.is_transpose and not conv2.is_transpose
            and conv1.in_channels % 4 == 0 and conv1.out_channels % 4 == 0
            and (downsample is None or (downsample.use_mm and downsample.bias is None))
            and x.F.is_cuda and x.F.size(0) > 0 and torch.is_grad_enabled())


def fused_residual_block(x, bn1, conv1, bn2, conv2, downsample=None):
    """common.py:21-50 as one call.  bn1/bn2: MinkowskiBatchNorm, conv1/conv2: MinkowskiConvolution(k=3),
    downsample: MinkowskiConvolution(k=1) or None.  Caller checks residual_block_fusable()."""
    mgr, key = x.coordinate_manager, x.coordinate_map_key
    kmap = mgr.kernel_map(key, key, 3)
    b1, b2 = bn1.bn, bn2.bn
    torch._foreach_add_([b1.num_batches_tracked, b2.num_batches_tracked], 1)
    out = _ResBlockFn.apply(x.F, b1.weight, b1.bias, conv1.kernel, b2.weight, b2.bias, conv2.kernel,
                            None if downsample is None else downsample.kernel, b1, b2, kmap,
                            conv1.packed(), conv2.packed(), None if downsample is None else downsample.packed())
    return SparseTensor(out, coordinate_map_key=key, coordinate_manager=mgr)


class _BnReluConvFn(torch.autograd.Function):
    """relu(bn(x)) -> strided convolution (mode 0) / transposed convolution (mode 1) of a U-Net level in one call
    (csrc/fused.cu: b2s_bnconv_forward/backward)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, kernel, bn, kmap, mode, packed=None):
        x = x.contiguous()
        n_x, cin = x.shape
        K, _, cout = kernel.shape
        dev = x.device
        n_out = kmap.n_out if mode == 0 else kmap.n_in
        saved = torch.empty(n_x * cin + 2 * cin, dtype=torch.float32, device=dev)  # y | stats
        out = torch.empty((n_out, cout), dtype=torch.float32, device=dev)
        p_y = saved.data_ptr()
        p_s = p_y + 4 * n_x * cin
        pin, pout, koff, maxp = kmap.pairs()
        algo = ops.get_conv_algo()
        c = max(cin, cout)
        ws = ops.workspace(ops.resblock_ws_bytes(K, c, c), dev)
        ops.check(ops.lib().b2s_bnconv_forward(
            x.data_ptr(), n_x, cin, cout, gamma.data_ptr(), beta.data_ptr(), ops.ptr(bn.running_mean),
            ops.ptr(bn.running_var), bn.eps, bn.momentum if bn.running_mean is not None else 0.0, kernel.data_ptr(),
            ops.ptr(packed), int(mode), kmap.nbr.data_ptr(), ops.ptr(kmap.tile_mask), pin.data_ptr(), pout.data_ptr(), koff.data_ptr(),
            int(maxp), kmap.n_out, kmap.n_in, K, p_y, p_s, out.data_ptr(), ops.bn_counter(dev).data_ptr(), algo,
            ws.data_ptr(), ws.numel(), ops.stream()), "bnconv_forward")
        ctx.save_for_backward(x, saved, gamma, kernel)
        ctx.kmap, ctx.mode, ctx.algo, ctx.packed = kmap, int(mode), algo, packed
        return out

    @staticmethod
    def backward(ctx, gout):
        x, saved, gamma, kernel = ctx.saved_tensors
        kmap, mode = ctx.kmap, ctx.mode
        gout = gout.contiguous()
        n_x, cin = x.shape
        K, _, cout = kernel.shape
        dev = x.device
        p_y = saved.data_ptr()
        p_s = p_y + 4 * n_x * cin
        gx = torch.empty_like(x)
        gw = torch.empty_like(kernel)
        dgb = torch.empty((2, cin), dtype=torch.float32, device=dev)
        tmp = torch.empty((n_x, cin), dtype=torch.float32, device=dev)
        pin, pout, koff, maxp = kmap.pairs()
        c = max(cin, cout)
        ws = ops.workspace(max(ops.resblock_ws_bytes(K, c, c), ops.wgrad_ws_bytes(K, cin, cout) + ops.lib().b2s_bn_ws_bytes(0, c) + 1024), dev)
        ops.check(ops.lib().b2s_bnconv_backward(
            gout.data_ptr(), x.data_ptr(), p_y, p_s, gamma.data_ptr(), kernel.data_ptr(), ops.ptr(ctx.packed), n_x, cin,
            cout, mode,
            kmap.nbr.data_ptr(), ops.ptr(kmap.tile_mask), pin.data_ptr(), pout.data_ptr(), koff.data_ptr(), int(maxp),
            kmap.n_out, kmap.n_in, K, gx.data_ptr(), gw.data_ptr(), dgb.data_ptr(), tmp.data_ptr(),
            ops.bn_counter(dev).data_ptr(), ctx.algo, ws.data_ptr(), ws.numel(), ops.stream()), "bnconv_backward")
        return gx, dgb[0], dgb[1], gw, None, None, None, None


def bn_relu_conv_fusable(x, bn_mod, conv):
    b = bn_mod.bn
    return (b.training and b.affine and b.track_running_stats and b.momentum is not None and conv.bias is None
            and conv.kernel_size == 2 and conv.stride == 2 and conv.in_channels % 4 == 0 and conv.out_channels % 4 == 0
            and x.F.is_cuda and x.F.size(0) > 0 and torch.is_grad_enabled())


def fused_bn_relu_conv(x, bn_mod, conv):
    """BN -> ReLU -> MinkowskiConvolution(k=2,s=2) or MinkowskiConvolutionTranspose(k=2,s=2) (common.py:67-77)."""
    mgr, key = x.coordinate_manager, x.coordinate_map_key
    b = bn_mod.bn
    if conv.is_transpose:
        out_key = mgr.existing_key(key.stride // conv.stride)
        kmap = mgr.kernel_map(out_key, key, conv.kernel_size)  # the forward strided map (fine -> coarse)
        mode = 1
    else:
        out_key = mgr.stride_key(key, conv.stride)
        kmap = mgr.kernel_map(key, out_key, conv.kernel_size)
        mode = 0
    b.num_batches_tracked += 1
    out = _BnReluConvFn.apply(x.F, b.weight, b.bias, conv.kernel, b, kmap, mode, conv.packed())
    return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)
