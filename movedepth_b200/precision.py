"""Arithmetic policy of the convolution GEMMs (cuDNN library calls) on the cost-volume branch.

The depth regression ends in an argmax (localmax, movedepth/layers.py:796-812), so `depth_mvs`
flips at individual pixels when the conv outputs move by more than a few ulp (SURVEY Appendix C5:
TF32 operands move 25 % of the pixels by >1e-3).  cuDNN's fp32 SIMT convolutions are exact enough
but 5-20x slower than its tensor-core kernels on B200, so the default policy is a **3xTF32 split**:

    x = x_hi + x_lo,  w = w_hi + w_lo     (x_hi = x rounded to TF32's 10-bit mantissa)
    conv(x, w) ~= conv(x_hi, w_hi) + conv(x_lo, w_hi) + conv(x_hi, w_lo)

evaluated as ONE tensor-core convolution over a 3x wider reduction dimension (channels for the
forward and the data gradient, batch for the weight gradient).  Every operand product is exact in
TF32; what remains is the tensor core's fp32 accumulation (~4e-6 relative, measured in
tools/bench_convs.py) instead of TF32's ~3e-4.

Policies: "fp32" (cuDNN SIMT kernels), "3xtf32" (above), "tf32" (PyTorch's default conv policy).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_policy = {"mode": "fp32"}


def set_policy(mode):
    assert mode in ("fp32", "3xtf32", "tf32"), mode
    _policy["mode"] = mode
    torch.backends.cudnn.allow_tf32 = mode != "fp32"


def get_policy():
    return _policy["mode"]


def tf32_round(x):
    """Round-to-nearest onto TF32's 10-bit mantissa (the low 13 bits become zero)."""
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _fmt(t):
    return torch.channels_last_3d if t.dim() == 5 else torch.channels_last


def _split(t, dim, pattern):
    """cat of (hi | lo) pieces of t along `dim`, e.g. pattern 'hlh' -> [hi, lo, hi]."""
    hi = tf32_round(t)
    lo = t - hi
    return torch.cat([hi if p == "h" else lo for p in pattern], dim)


class _SplitConv(torch.autograd.Function):
    """y = conv(x, w) (+ transposed variant) with the 3xTF32 split in forward, dgrad and wgrad."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, output_padding, transposed):
        x = x.contiguous(memory_format=_fmt(x))
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, output_padding, transposed)
        x3 = _split(x, 1, "hlh")
        w3 = _split(w, 0 if transposed else 1, "hhl").contiguous(memory_format=_fmt(w))
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
            if ctx.needs_input_grad[0]:
                # reduction over output channels -> split along them
                gy3 = _split(gy, 1, "hlh")
                w3 = _split(w, 1 if transposed else 0, "hhl").contiguous(memory_format=_fmt(w))
                gx = torch.ops.aten.convolution_backward(gy3, x, w3, None, stride, padding, (1,) * nd, transposed,
                                                         output_padding, 1, [True, False, False])[0]
            if ctx.needs_input_grad[1]:
                # reduction over batch and positions -> split along the batch
                xb = _split(x, 0, "hlh")
                gb = _split(gy, 0, "hhl")
                gw = torch.ops.aten.convolution_backward(gb, xb, w, None, stride, padding, (1,) * nd, transposed,
                                                         output_padding, 1, [False, True, False])[1]
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return gx, gw, None, None, None, None


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class _PolicyMixin:
    _transposed = False

    def forward(self, x):
        mode = _policy["mode"]
        if mode != "3xtf32":
            return super().forward(x)
        nd = x.dim() - 2
        y = _SplitConv.apply(x, self.weight, _tup(self.stride, nd), _tup(self.padding, nd),
                             _tup(getattr(self, "output_padding", 0), nd), self._transposed)
        if self.bias is not None:
            y = y + self.bias.view(1, -1, *([1] * nd))
        return y


class Conv2d(_PolicyMixin, nn.Conv2d):
    """nn.Conv2d (same parameters / state-dict keys) that honours the branch precision policy."""


class Conv3d(_PolicyMixin, nn.Conv3d):
    """nn.Conv3d that honours the branch precision policy."""


class ConvTranspose3d(_PolicyMixin, nn.ConvTranspose3d):
    """nn.ConvTranspose3d that honours the branch precision policy."""
    _transposed = True
