"""Arithmetic policy of the convolution GEMMs (cuDNN library calls).

The reference is fp32.  Two of its outputs are precision-sensitive: the mono disparity feeds the
hypothesis range, and the depth regression ends in an argmax (localmax, movedepth/layers.py:796-812),
so `depth_mvs` flips at individual pixels when conv outputs move by more than a few ulp (SURVEY
Appendix C5; measured here in tools/parity_report.py: plain TF32 leaves only ~62 % of the pixels
within 1e-3 of the reference).  cuDNN's fp32 SIMT convolutions are exact enough but 5-20x slower
than its tensor-core kernels on B200, so the default policy is a **3xTF32 split** of the forward:

    x = x_hi + x_lo,  w = w_hi + w_lo     (x_hi = x rounded to TF32's 10-bit mantissa)
    conv(x, w) ~= conv(x_hi, w_hi) + conv(x_lo, w_hi) + conv(x_hi, w_lo)

evaluated as ONE tensor-core convolution over a 3x wider channel dimension.  Every operand product
is exact in TF32; what remains is the tensor core's fp32 accumulation (~4e-6 relative, measured in
tools/bench_convs.py) instead of TF32's ~3e-4.  The operand split is one hand-written kernel
(`mvd_split_tf32`).  Gradients (dgrad / wgrad) use single-pass TF32 -- PyTorch's default conv
policy on this hardware -- unless `split_backward` is set.

Policies: "fp32" (cuDNN SIMT kernels, bit-level parity runs), "3xtf32" (default), "tf32".
"""
import ctypes

import torch
import torch.nn as nn

_policy = {"mode": "fp32", "split_backward": False}
conv_recorder = None     # bench.py's step model: when a list, code paths that bypass nn.Module.__call__ of a conv append
                         # (weight shape, input shape, output shape, transposed) here so that every convolution is counted


def record_conv(weight, x, y, transposed=False):
    if conv_recorder is not None:
        conv_recorder.append((tuple(weight.shape), tuple(x.shape), tuple(y.shape), bool(transposed)))


small_direct = True      # skinny 2-D layers (<= 16 channels) on the exact-fp32 direct kernels of csrc/conv2d_small.cu


def set_policy(mode, split_backward=None):
    assert mode in ("fp32", "3xtf32", "tf32"), mode
    _policy["mode"] = mode
    if split_backward is not None:
        _policy["split_backward"] = bool(split_backward)
    torch.backends.cudnn.allow_tf32 = mode != "fp32"


def get_policy():
    return _policy["mode"]


def tf32_round(x):
    """Round-to-nearest onto TF32's 10-bit mantissa (the low 13 bits become zero)."""
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _fmt(t):
    return torch.channels_last_3d if t.dim() == 5 else torch.channels_last


def _split_dim1(t, weight):
    """[N, C, ...] channels-last -> [N, 3C, ...] channels-last holding [hi, lo, hi] (activations) or
    [hi, hi, lo] (weights) along the channel dimension."""
    t = t.contiguous(memory_format=_fmt(t))
    C = t.shape[1]
    if t.is_cuda:
        from . import _lib, ops
        out = torch.empty((t.shape[0], 3 * C) + tuple(t.shape[2:]), device=t.device, dtype=torch.float32,
                          memory_format=_fmt(t))
        rc = _lib.lib().mvd_split_tf32(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(out.data_ptr()), t.numel(), C,
                                       1 if weight else 0, ops._stream())
        _lib.check(rc, "mvd_split_tf32")
        ops.launch_counter["n"] += 1
        return out
    hi = tf32_round(t)
    lo = t - hi
    return torch.cat([hi, hi, lo] if weight else [hi, lo, hi], 1)


def _split_dim0(t, weight):
    hi = tf32_round(t)
    lo = t - hi
    return torch.cat([hi, hi, lo] if weight else [hi, lo, hi], 0).contiguous(memory_format=_fmt(t))


class _SplitConv(torch.autograd.Function):
    """y = conv(x, w) (or its transposed variant) with the 3xTF32 operand split."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, output_padding, transposed):
        x = x.contiguous(memory_format=_fmt(x))
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, output_padding, transposed)
        x3 = _split_dim1(x, False)
        w3 = _split_dim0(w, True) if transposed else _split_dim1(w, True)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            nd = x.dim() - 2
            y = torch.ops.aten.convolution(x3, w3, None, stride, padding, (1,) * nd, transposed, output_padding, 1)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, output_padding, transposed = ctx.cfg
        nd = x.dim() - 2
        gy = gy.contiguous(memory_format=_fmt(gy))
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        gx = gw = None
        try:
            need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            if (nd == 3 and not transposed and x.is_cuda and tuple(w.shape) == (16, 16, 3, 3, 3) and tuple(stride) == (1, 1, 1)
                    and tuple(padding) == (1, 1, 1) and not _policy["split_backward"]):
                # reg3d's full-resolution 16->16 layer: hand-written gradients (cuDNN needs 1.6 ms + 2.6 ms here)
                from . import ops
                gx = ops.c16c16_conv_tc(gy, w, 1, 1) if need_x else None      # tcgen05 / TMEM implicit GEMM, TF32
                gw = ops.c16c16_wgrad_tc(gy, x) if need_w else None                   # tcgen05, MN-major operands, TF32
            elif not _policy["split_backward"]:
                gx, gw, _ = torch.ops.aten.convolution_backward(gy, x, w.contiguous(memory_format=_fmt(w)), None, stride,
                                                                padding, (1,) * nd, transposed, output_padding, 1,
                                                                [need_x, need_w, False])
            else:
                if need_x:       # reduction over output channels -> split along them
                    gy3 = _split_dim1(gy, False)
                    w3 = _split_dim1(w, True) if transposed else _split_dim0(w, True)
                    gx = torch.ops.aten.convolution_backward(gy3, x, w3, None, stride, padding, (1,) * nd, transposed,
                                                             output_padding, 1, [True, False, False])[0]
                if need_w:       # reduction over batch and positions -> split along the batch
                    xb = _split_dim0(x, False)
                    gb = _split_dim0(gy, True)
                    gw = torch.ops.aten.convolution_backward(gb, xb, w, None, stride, padding, (1,) * nd, transposed,
                                                             output_padding, 1, [False, True, False])[1]
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return gx, gw, None, None, None, None


class _PreparedConv(torch.autograd.Function):
    """y = conv2d(xp, w) with NO padding on an input that is already padded (and, under 3xTF32, already split: `x3`)."""

    @staticmethod
    def forward(ctx, xp, x3, w):
        ctx.save_for_backward(xp, w)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            return torch.ops.aten.convolution(x3, _split_dim1(w, True), None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1)
        finally:
            torch.backends.cudnn.allow_tf32 = prev

    @staticmethod
    def backward(ctx, gy):
        xp, w = ctx.saved_tensors
        gy = gy.contiguous(memory_format=torch.channels_last)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            gx, gw, _ = torch.ops.aten.convolution_backward(gy, xp, w.contiguous(memory_format=torch.channels_last), None, (1, 1),
                                                            (0, 0), (1, 1), False, (0, 0), 1,
                                                            [ctx.needs_input_grad[0], ctx.needs_input_grad[2], False])
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return gx, None, gw


def prepared_is_small(weight):
    """True if `prepared_conv` runs this layer on the exact-fp32 direct kernels (csrc/conv2d_small.cu): the DepthDecoder's finest
    stage, 16 -> 16 and the 16 -> 1 disparity head (no operand split is needed for its input then)."""
    return (small_direct and weight.is_cuda and weight.dtype == torch.float32 and weight.dim() == 4 and tuple(weight.shape[2:]) == (3, 3)
            and weight.shape[1] == 16 and weight.shape[0] in (16, 1))


def prepared_conv(xp, x3, weight):
    """3x3 (or any) stride-1 convolution, no padding, no bias, of a pre-padded channels-last input following the
    precision policy; `x3` is the pre-split operand the DepthDecoder glue kernel wrote (used under 3xtf32 only)."""
    if prepared_is_small(weight) and xp.is_cuda:
        from . import ops
        w = weight
        if w.shape[0] == 1:                       # the one-channel head rides the 4-output-channel kernel (zero filters 1..3)
            w = torch.cat([w, w.new_zeros((3,) + tuple(w.shape[1:]))], 0)
        y = ops.conv2d_small(xp, w, 3, 1, pad=0)
        if weight.shape[0] == 1:
            y = y[:, :1]
        record_conv(weight, xp, y)
        return y
    if _policy["mode"] == "3xtf32" and not _policy["split_backward"] and x3.numel():
        y = _PreparedConv.apply(xp, x3, weight)
    elif _policy["mode"] == "3xtf32":
        y = _SplitConv.apply(xp, weight, (1, 1), (0, 0), (0, 0), False)
    else:
        y = torch.nn.functional.conv2d(xp, weight)
    record_conv(weight, xp, y)
    return y


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class _PolicyMixin:
    _transposed = False

    def _small(self, x):
        """(k, stride) when this is one of the skinny 2-D layers the direct fp32 kernels cover, else None."""
        if not (small_direct and x.is_cuda and x.dim() == 4 and not self._transposed and self.bias is None and self.groups == 1
                and x.dtype == torch.float32 and getattr(self, "padding_mode", "zeros") == "zeros"):
            return None
        k, s, p, d = _tup(self.kernel_size, 2), _tup(self.stride, 2), _tup(self.padding, 2), _tup(self.dilation, 2)
        if k[0] != k[1] or s[0] != s[1] or p != (k[0] // 2, k[0] // 2) or d != (1, 1):
            return None
        from . import ops
        return (k[0], s[0]) if ops.conv2d_small_supported(self.in_channels, self.out_channels, k[0], s[0]) else None

    def forward(self, x):
        small = self._small(x)
        if small is not None and (self.in_channels > 3 or not x.requires_grad):      # the 3-channel layer has no data gradient
            from . import ops
            return ops.conv2d_small(x, self.weight, *small)
        if _policy["mode"] != "3xtf32" or getattr(self, "padding_mode", "zeros") != "zeros" or self.groups != 1 \
                or any(d != 1 for d in _tup(self.dilation, x.dim() - 2)):
            return super().forward(x)
        nd = x.dim() - 2
        if (nd == 3 and not self._transposed and x.is_cuda and self.bias is None and tuple(self.weight.shape) == (16, 16, 3, 3, 3)
                and _tup(self.stride, 3) == (1, 1, 1) and _tup(self.padding, 3) == (1, 1, 1) and not _policy["split_backward"]):
            from . import ops                 # reg3d's full-resolution 16->16 layer: hand-written tcgen05 implicit GEMM
            return ops.conv3d_c16_to_16(x, self.weight, 3)
        y = _SplitConv.apply(x, self.weight, _tup(self.stride, nd), _tup(self.padding, nd),
                             _tup(getattr(self, "output_padding", 0), nd), self._transposed)
        if self.bias is not None:
            y = y + self.bias.view(1, -1, *([1] * nd))
        return y


class Conv2d(_PolicyMixin, nn.Conv2d):
    """nn.Conv2d (same parameters / state-dict keys) that honours the precision policy."""


class Conv3d(_PolicyMixin, nn.Conv3d):
    """nn.Conv3d that honours the precision policy."""


class ConvTranspose3d(_PolicyMixin, nn.ConvTranspose3d):
    """nn.ConvTranspose3d that honours the precision policy."""
    _transposed = True


_SWAP = {nn.Conv2d: Conv2d, nn.Conv3d: Conv3d, nn.ConvTranspose3d: ConvTranspose3d}


def adopt(module):
    """Make every plain conv inside `module` (e.g. a torchvision ResNet) policy-aware, in place.
    Parameters, buffers and state-dict keys are untouched (only the class of the conv modules changes)."""
    for m in module.modules():
        cls = _SWAP.get(type(m))
        if cls is not None:
            m.__class__ = cls
    return module
