"""torch.autograd bindings of the hand-written sm_100a kernels (through the C ABI, via ctypes).

PyTorch is used here for device memory, streams and the autograd tape only; all arithmetic of
these ops happens in libmovedepth_b200.so.  CPU tensors are rejected: there is no fallback.
"""
import ctypes

import torch

from . import _lib

LAYOUT_BGDHW = 0
LAYOUT_BDHWG = 1
FLAG_NO_TMA = 1
FLAG_NO_TABLE = 2
FLAG_BWD_V2 = 32        # backward: the round-1 kernel (A/B runs)

# counts kernel launches issued through the C ABI (bench.py reports it as `gpu_launches`)
launch_counter = {"n": 0}
# when set to a list, the cost-volume forward appends (start, end) CUDA events recorded on the
# launching stream around its kernel (bench.py's live roofline measurement)
costvol_events = None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class TimingEvent:
    """CUDA timing event that can also be recorded inside a captured graph (cudaEventRecordExternal): the record
    then fires at every replay, and `elapsed_ms` reads the pair after a synchronize."""

    def __init__(self):
        self.handle = ctypes.c_void_p(_lib.lib().mvd_event_create())
        if not self.handle:
            raise RuntimeError("cudaEventCreate failed")

    def record(self):
        _lib.check(_lib.lib().mvd_event_record(self.handle, _stream(), 1 if torch.cuda.is_current_stream_capturing() else 0),
                   "mvd_event_record")

    def elapsed_ms(self, stop):
        ms = ctypes.c_float(0)
        _lib.check(_lib.lib().mvd_event_elapsed_ms(self.handle, stop.handle, ctypes.byref(ms)), "mvd_event_elapsed_ms")
        return ms.value

    def elapsed_time(self, stop):            # torch.cuda.Event spelling
        return self.elapsed_ms(stop)


def _p(t):
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("movedepth_b200 ops need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def _mat(m, B):
    m = _f32(m).reshape(-1, 4, 4)
    if m.shape[0] == 1 and B > 1:
        m = m.expand(B, 4, 4)
    assert m.shape[0] == B, "expected %d 4x4 matrices, got %d" % (B, m.shape[0])
    return m.contiguous()


def _nhwc(t):
    return _f32(t).contiguous(memory_format=torch.channels_last)


# ------------------------------------------------------------------------------------- K1
class _CostVolumeGrouped(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, src, prior, ratio, hyps, K, invK, T, groups, layout, flags):
        B, C, h, w = ref.shape
        ref_cl, src_cl = _nhwc(ref), _nhwc(src)
        if hyps is not None:
            hyps = _f32(hyps).contiguous()
            D = hyps.shape[1]
            prior_c = ratio_c = None
        else:
            prior_c = _f32(prior).reshape(B, h, w).contiguous()
            ratio_c = _f32(ratio).contiguous()
            D = ratio_c.shape[1]
        K, invK, T = _mat(K, B), _mat(invK, B), _mat(T, B)
        fmt = torch.contiguous_format if layout == LAYOUT_BGDHW else torch.channels_last_3d
        out = torch.empty((B, groups, D, h, w), device=ref.device, dtype=torch.float32, memory_format=fmt)
        ev = None
        if costvol_events is not None:
            ev = (TimingEvent(), TimingEvent())
            ev[0].record()
        rc = _lib.lib().mvd_costvol_grouped_fwd(_p(ref_cl), _p(src_cl), _p(prior_c), _p(ratio_c), _p(hyps), _p(K),
                                                _p(invK), _p(T), _p(out), B, C, groups, h, w, D, layout, flags, _stream())
        _lib.check(rc, "mvd_costvol_grouped_fwd")
        if ev is not None:
            ev[1].record()
            costvol_events.append(ev)
        launch_counter["n"] += 1
        ctx.save_for_backward(ref_cl, src_cl, prior_c, ratio_c, hyps, K, invK, T)
        ctx.meta = (B, C, groups, h, w, D, layout, flags, fmt)
        return out

    @staticmethod
    def backward(ctx, gout):
        ref_cl, src_cl, prior_c, ratio_c, hyps, K, invK, T = ctx.saved_tensors
        B, C, groups, h, w, D, layout, flags, fmt = ctx.meta
        gout = _f32(gout).contiguous(memory_format=fmt)
        both = torch.empty((2, B, h, w, C), device=ref_cl.device, dtype=ref_cl.dtype).permute(0, 1, 4, 2, 3)
        gref, gsrc = both[0], both[1]                              # channels-last like ref / src; adjacent: one memset in the library
        rc = _lib.lib().mvd_costvol_grouped_bwd(_p(gout), _p(ref_cl), _p(src_cl), _p(prior_c), _p(ratio_c), _p(hyps),
                                                _p(K), _p(invK), _p(T), _p(gref), _p(gsrc), B, C, groups, h, w, D,
                                                layout, flags, _stream())
        _lib.check(rc, "mvd_costvol_grouped_bwd")
        launch_counter["n"] += 2
        return gref, gsrc, None, None, None, None, None, None, None, None, None


def costvol_grouped(ref, src, K, invK, T, prior=None, ratio=None, hyps=None, groups=16, layout=LAYOUT_BGDHW, flags=0):
    """Fused warp + gather + group correlation.  Returns the grouped volume as a logical
    [B,G,D,h,w] tensor (physically [B,D,h,w,G] when layout == LAYOUT_BDHWG).
    Reference: movedepth/layers.py:778-794 + movedepth/trainer.py:359."""
    return _CostVolumeGrouped.apply(ref, src, prior, ratio, hyps, K, invK, T, groups, layout, flags)


class _CostVolumeFull(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, src, hyps, K, invK, T):
        B, C, h, w = ref.shape
        ref, src, hyps = _f32(ref).contiguous(), _f32(src).contiguous(), _f32(hyps).contiguous()
        D = hyps.shape[1]
        K, invK, T = _mat(K, B), _mat(invK, B), _mat(T, B)
        out = torch.empty((B, D, C, h, w), device=ref.device, dtype=torch.float32)
        rc = _lib.lib().mvd_costvol_full_fwd(_p(ref), _p(src), _p(hyps), _p(K), _p(invK), _p(T), _p(out), B, C, h, w, D,
                                             _stream())
        _lib.check(rc, "mvd_costvol_full_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(ref, src, hyps, K, invK, T)
        return out

    @staticmethod
    def backward(ctx, gout):
        ref, src, hyps, K, invK, T = ctx.saved_tensors
        B, C, h, w = ref.shape
        D = hyps.shape[1]
        gout = _f32(gout).contiguous()
        gref, gsrc = torch.empty_like(ref), torch.empty_like(src)
        rc = _lib.lib().mvd_costvol_full_bwd(_p(gout), _p(ref), _p(src), _p(hyps), _p(K), _p(invK), _p(T), _p(gref),
                                             _p(gsrc), B, C, h, w, D, _stream())
        _lib.check(rc, "mvd_costvol_full_bwd")
        launch_counter["n"] += 3
        return gref, gsrc, None, None, None, None


def costvol_full(ref, src, hyps, K, invK, T):
    """Reference-layout volume [B,D,C,h,w] (movedepth/layers.py:778-794)."""
    return _CostVolumeFull.apply(ref, src, hyps, K, invK, T)


# ------------------------------------------------------------------------------------- K3
class _Regress(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, inv_a, inv_b, radius, want_prob):
        B, D, h, w = logits.shape
        logits = _f32(logits).contiguous()
        inv_a = _f32(inv_a).reshape(B, h, w).contiguous()
        inv_b = _f32(inv_b).reshape(B, h, w).contiguous()
        prob = torch.empty_like(logits) if want_prob else None
        ent = torch.empty((B, 1, h, w), device=logits.device, dtype=torch.float32)
        depth = torch.empty((B, h, w), device=logits.device, dtype=torch.float32)
        amax = torch.empty((B, h, w), device=logits.device, dtype=torch.int32)
        rc = _lib.lib().mvd_regress_fwd(_p(logits), _p(inv_a), _p(inv_b), _p(prob), _p(ent), _p(depth), _p(amax), B, D,
                                        h * w, radius, _stream())
        _lib.check(rc, "mvd_regress_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(logits, inv_a, inv_b, amax)
        ctx.radius = radius
        if prob is None:
            prob = logits.new_empty(0)
        ctx.mark_non_differentiable(prob)
        return prob, ent, depth

    @staticmethod
    def backward(ctx, _gprob, gent, gdepth):
        logits, inv_a, inv_b, amax = ctx.saved_tensors
        B, D, h, w = logits.shape
        gent = None if gent is None else _f32(gent).contiguous()
        gdepth = None if gdepth is None else _f32(gdepth).contiguous()
        gl = torch.empty_like(logits)
        rc = _lib.lib().mvd_regress_bwd(_p(logits), _p(inv_a), _p(inv_b), _p(amax), _p(gent), _p(gdepth), _p(gl), B, D,
                                        h * w, ctx.radius, _stream())
        _lib.check(rc, "mvd_regress_bwd")
        launch_counter["n"] += 1
        return gl, None, None, None, None


def regress_depth(logits, inv_a, inv_b, radius=1, want_prob=False):
    """softmax over D + entropy + local-max regression in one kernel.
    Returns (prob [B,D,h,w] or empty, entropy [B,1,h,w], depth [B,h,w]).  prob carries no gradient.
    Reference: movedepth/trainer.py:367-371, layers.py:796-812, 862-863."""
    return _Regress.apply(logits, inv_a, inv_b, radius, want_prob)


# ------------------------------------------------------------------------------------- K4
class _ConvexUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, mask, f):
        B, h, w = depth.shape
        depth, mask = _f32(depth).contiguous(), _f32(mask).contiguous()
        assert mask.shape == (B, 9 * f * f, h, w), mask.shape
        out = torch.empty((B, f * h, f * w), device=depth.device, dtype=torch.float32)
        rc = _lib.lib().mvd_convex_up_fwd(_p(depth), _p(mask), _p(out), B, h, w, f, _stream())
        _lib.check(rc, "mvd_convex_up_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(depth, mask)
        ctx.f = f
        return out

    @staticmethod
    def backward(ctx, gout):
        depth, mask = ctx.saved_tensors
        B, h, w = depth.shape
        gout = _f32(gout).contiguous()
        gdepth, gmask = torch.empty_like(depth), torch.empty_like(mask)
        rc = _lib.lib().mvd_convex_up_bwd(_p(depth), _p(mask), _p(gout), _p(gdepth), _p(gmask), B, h, w, ctx.f, _stream())
        _lib.check(rc, "mvd_convex_up_bwd")
        launch_counter["n"] += 2
        return gdepth, gmask, None


def convex_upsample(depth, mask, scale=2):
    """movedepth/layers.py:200-214.  depth [B,h,w] or [B,1,h,w]; mask [B,9*4**scale,h,w]."""
    if depth.dim() == 4:
        depth = depth[:, 0]
    return _ConvexUp.apply(depth, mask, 2 ** scale)


# ------------------------------------------------------------------------------------- loss assembly
class _ReprojSelect(torch.autograd.Function):
    @staticmethod
    def forward(ctx, l0, l1, ident, noise):
        l0 = _f32(l0).contiguous()
        n = l0.numel()
        l1c = _f32(l1).contiguous() if l1 is not None else None
        identc = _f32(ident.detach()).contiguous() if ident is not None else None
        noisec = _f32(noise.detach()).contiguous() if noise is not None else None
        reproj = torch.empty_like(l0)
        sel = torch.empty(n, device=l0.device, dtype=torch.uint8)
        sums = torch.empty(2, device=l0.device, dtype=torch.float64)
        loss = torch.empty((), device=l0.device, dtype=torch.float32)
        rc = _lib.lib().mvd_reproj_select_fwd(_p(l0), _p(l1c), _p(identc), _p(noisec), _p(reproj), _p(sel), _p(sums), _p(loss), n,
                                              _stream())
        _lib.check(rc, "mvd_reproj_select_fwd")
        launch_counter["n"] += 3
        ctx.save_for_backward(sel, sums)
        ctx.shape, ctx.two = l0.shape, l1 is not None
        ctx.mark_non_differentiable(reproj)
        return loss, reproj

    @staticmethod
    def backward(ctx, gloss, _greproj):
        sel, sums = ctx.saved_tensors
        gloss = _f32(gloss).contiguous()
        g0 = torch.empty(ctx.shape, device=sel.device, dtype=torch.float32)
        g1 = torch.empty(ctx.shape, device=sel.device, dtype=torch.float32) if ctx.two else None
        rc = _lib.lib().mvd_reproj_select_bwd(_p(gloss), _p(sums), _p(sel), _p(g0), _p(g1), sel.numel(), _stream())
        _lib.check(rc, "mvd_reproj_select_bwd")
        launch_counter["n"] += 1
        return g0, g1, None, None


def reproj_select(per_src, ident=None, noise=None):
    """min over the source frames' reprojection losses, identity auto-mask, masked mean -- one kernel.
    per_src: list of 1 or 2 [B,1,H,W] loss maps; ident: [B,1,H,W] identity loss (already min over sources) or None
    (mask = ones); noise: N(0,1) tie-break noise or None.  Returns (loss scalar, reproj [B,1,H,W]).
    Reference: movedepth/trainer.py:687-709, 621-662, 589-609."""
    assert 1 <= len(per_src) <= 2
    return _ReprojSelect.apply(per_src[0], per_src[1] if len(per_src) == 2 else None, ident, noise)


# ------------------------------------------------------------------------------------- reg3d output head
class _Conv3dC16O1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        B, C, D, H, W = x.shape
        assert C == 16 and tuple(weight.shape) == (1, 16, 3, 3, 3), (x.shape, weight.shape)
        x = _f32(x).contiguous(memory_format=torch.channels_last_3d)
        w = _f32(weight).contiguous()
        y = torch.empty((B, 1, D, H, W), device=x.device, dtype=torch.float32)
        rc = _lib.lib().mvd_conv3d_c16o1_fwd(_p(x), _p(w), _p(y), B, D, H, W, _stream())
        _lib.check(rc, "mvd_conv3d_c16o1_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        B, _, D, H, W = x.shape
        gy = _f32(gy).contiguous()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)                      # channels-last-3d
            rc = _lib.lib().mvd_conv3d_c16o1_dgrad(_p(gy), _p(w), _p(gx), B, D, H, W, _stream())
            _lib.check(rc, "mvd_conv3d_c16o1_dgrad")
            launch_counter["n"] += 1
        if ctx.needs_input_grad[1]:
            nbytes = _lib.lib().mvd_conv3d_c16o1_wgrad_workspace_bytes(B, D, H, W)
            ws = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
            gw = torch.empty_like(w)
            rc = _lib.lib().mvd_conv3d_c16o1_wgrad(_p(gy), _p(x), _p(gw), _p(ws), nbytes, B, D, H, W, _stream())
            _lib.check(rc, "mvd_conv3d_c16o1_wgrad")
            launch_counter["n"] += 2
        return gx, gw


def conv3d_c16_to_1(x, weight):
    """Conv3d(16 -> 1, k=3, stride 1, padding 1, no bias) on a channels-last-3d volume, exact fp32.
    x: logical [B,16,D,H,W]; weight: [1,16,3,3,3]; returns [B,1,D,H,W].
    Reference: reg3d.prob, movedepth/networks/resnet_encoder.py:254, 279."""
    return _Conv3dC16O1.apply(x, weight)


# ------------------------------------------------------------------------------------- reg3d first layer
def c16c16_conv(inp, weight, mode, passes):
    """Raw launch: mode 0 = forward, 1 = data gradient; inp logical [B,16,D,H,W] channels-last-3d, weight [16,16,3,3,3]."""
    B, C, D, H, W = inp.shape
    assert C == 16 and tuple(weight.shape) == (16, 16, 3, 3, 3), (inp.shape, weight.shape)
    inp = _f32(inp).contiguous(memory_format=torch.channels_last_3d)
    w = _f32(weight).contiguous()
    out = torch.empty_like(inp)
    rc = _lib.lib().mvd_conv3d_c16c16(_p(inp), _p(w), _p(out), B, D, H, W, mode, passes, _stream())
    _lib.check(rc, "mvd_conv3d_c16c16")
    launch_counter["n"] += 1
    return out


def c16c16_conv_tc(inp, weight, mode, passes, flags=0, bn_sums=None):
    """Same contract as c16c16_conv on tcgen05 / TMEM (csrc/conv3d_c16.cu, namespace tc).  `bn_sums` (zeroed fp64 [>= 32]):
    the epilogue accumulates the output's per-channel sum and sum of squares into it (fused BatchNorm statistics)."""
    B, C, D, H, W = inp.shape
    assert C == 16 and tuple(weight.shape) == (16, 16, 3, 3, 3), (inp.shape, weight.shape)
    inp = _f32(inp).contiguous(memory_format=torch.channels_last_3d)
    w = _f32(weight).contiguous()
    out = torch.empty_like(inp)
    rc = _lib.lib().mvd_conv3d_c16c16_tc(_p(inp), _p(w), _p(out), _p(bn_sums), B, D, H, W, mode, passes, flags, _stream())
    _lib.check(rc, "mvd_conv3d_c16c16_tc")
    launch_counter["n"] += 1
    return out


def c16c16_wgrad(gy, x):
    """Exact-fp32 weight gradient [16,16,3,3,3] of the 16->16 layer; gy, x logical [B,16,D,H,W] channels-last-3d."""
    B, C, D, H, W = x.shape
    gy = _f32(gy).contiguous(memory_format=torch.channels_last_3d)
    x = _f32(x).contiguous(memory_format=torch.channels_last_3d)
    nbytes = _lib.lib().mvd_conv3d_c16c16_wgrad_workspace_bytes(B, D, H, W)
    ws = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
    gw = torch.empty((16, 16, 3, 3, 3), device=x.device, dtype=torch.float32)
    rc = _lib.lib().mvd_conv3d_c16c16_wgrad(_p(gy), _p(x), _p(gw), _p(ws), nbytes, B, D, H, W, _stream())
    _lib.check(rc, "mvd_conv3d_c16c16_wgrad")
    launch_counter["n"] += 2
    return gw


def c16c16_wgrad_tc(gy, x):
    """TF32 weight gradient [16,16,3,3,3] of the 16->16 layer on tcgen05 / TMEM (odd widths: exact-fp32 FFMA2 kernel)."""
    B, C, D, H, W = x.shape
    if W % 2:
        return c16c16_wgrad(gy, x)
    gy = _f32(gy).contiguous(memory_format=torch.channels_last_3d)
    x = _f32(x).contiguous(memory_format=torch.channels_last_3d)
    nbytes = _lib.lib().mvd_conv3d_c16c16_wgrad_tc_workspace_bytes(B, D, H, W)
    ws = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
    gw = torch.empty((16, 16, 3, 3, 3), device=x.device, dtype=torch.float32)
    rc = _lib.lib().mvd_conv3d_c16c16_wgrad_tc(_p(gy), _p(x), _p(gw), _p(ws), nbytes, B, D, H, W, _stream())
    _lib.check(rc, "mvd_conv3d_c16c16_wgrad_tc")
    launch_counter["n"] += 2
    return gw


class _Conv3dC16C16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, passes, want_stats):
        sums = torch.zeros(33, device=x.device, dtype=torch.float64) if want_stats else None   # [sum, sum sq, BN arrival counter]
        y = c16c16_conv_tc(x, weight, 0, passes, bn_sums=sums)                  # tcgen05 / TMEM implicit GEMM
        if want_stats:
            launch_counter["n"] += 1                                            # the zero fill
        ctx.save_for_backward(x.contiguous(memory_format=torch.channels_last_3d), weight)
        if sums is None:
            sums = y.new_empty(0, dtype=torch.float64)
        ctx.mark_non_differentiable(sums)
        return y, sums

    @staticmethod
    def backward(ctx, gy, _gsums):
        x, w = ctx.saved_tensors
        gx = c16c16_conv_tc(gy, w, 1, 1) if ctx.needs_input_grad[0] else None   # single-pass TF32 data gradient (tcgen05)
        gw = c16c16_wgrad_tc(gy, x) if ctx.needs_input_grad[1] else None        # single-pass TF32 weight gradient (tcgen05)
        return gx, gw, None, None


def conv3d_c16_to_16(x, weight, passes=3):
    """Conv3d(16 -> 16, k=3, stride 1, padding 1, no bias) on a channels-last-3d volume, all three passes hand-written:
    tcgen05/TMEM implicit GEMM forward (3xTF32 split with passes=3, plain TF32 with passes=1), TF32 data gradient and
    TF32 weight gradient (MN-major operands); `c16c16_wgrad` is the exact-fp32 FFMA2 weight gradient.  Reference: reg3d.conv0.conv, movedepth/networks/resnet_encoder.py:178, 231."""
    return _Conv3dC16C16.apply(x, weight, passes, False)[0]


def conv3d_c16_to_16_with_stats(x, weight, passes=3):
    """conv3d_c16_to_16 whose epilogue also accumulates the BatchNorm statistics of its output: returns (y, sums) with
    sums = fp64 [sum y (16), sum y^2 (16), 0] ready for `norm.bn_act(..., sums=sums)` (no separate statistics pass)."""
    return _Conv3dC16C16.apply(x, weight, passes, True)


# ------------------------------------------------------------------------------------- Adam
def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """One fused Adam update over flat fp32 arenas (torch.optim.Adam semantics)."""
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    step_size = lr / (1.0 - beta1 ** step)
    bias2 = (1.0 - beta2 ** step) ** 0.5
    rc = _lib.lib().mvd_adam_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), n, beta1, beta2, eps, step_size,
                                  bias2, grad_scale, _stream())
    _lib.check(rc, "mvd_adam_step")
    launch_counter["n"] += 1


# ------------------------------------------------------------------------------------- K5 + K6
class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, src, tgt, K, invK, T, ssim_w):
        B, _, H, W = src.shape
        depth = _f32(depth).reshape(B, H, W).contiguous()
        src, tgt = _f32(src).contiguous(), _f32(tgt).contiguous()
        K, invK, Tm = _mat(K, B), _mat(invK, B), _mat(T, B)
        warped = torch.empty_like(src)
        loss = torch.empty((B, 1, H, W), device=src.device, dtype=torch.float32)
        rc = _lib.lib().mvd_photometric_fwd(_p(depth), _p(src), _p(tgt), _p(K), _p(invK), _p(Tm), _p(warped), _p(loss), B, H,
                                            W, float(ssim_w), 0, _stream())
        _lib.check(rc, "mvd_photometric_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(depth, src, tgt, warped, K, invK, Tm)
        ctx.ssim_w = float(ssim_w)
        ctx.depth_shape = None
        ctx.mark_non_differentiable(warped)
        return loss, warped

    @staticmethod
    def backward(ctx, gloss, _gwarped):
        depth, src, tgt, warped, K, invK, Tm = ctx.saved_tensors
        B, _, H, W = src.shape
        gloss = _f32(gloss).contiguous()
        need_T = ctx.needs_input_grad[5]
        gdepth = torch.empty_like(depth)
        gP = torch.empty((B, 3, 4), device=src.device, dtype=torch.float32) if need_T else None
        rc = _lib.lib().mvd_photometric_bwd(_p(gloss), _p(depth), _p(src), _p(tgt), _p(warped), _p(K), _p(invK), _p(Tm),
                                            _p(gdepth), _p(gP), B, H, W, ctx.ssim_w, _stream())
        _lib.check(rc, "mvd_photometric_bwd")
        launch_counter["n"] += 2 if need_T else 1
        gT = None
        if need_T:                                   # P = K @ T  ->  dL/dT = K^T[:, :3] @ dL/dP  (4x3 @ 3x4, glue)
            gT = torch.matmul(K[:, :3, :].transpose(1, 2), gP)
        return gdepth, None, None, None, None, gT, None


def photometric_loss(depth, src, tgt, K, invK, T, ssim_w=0.85):
    """Warp `src` into the target view with per-pixel `depth` and compare with `tgt`.
    Returns (loss [B,1,H,W], warped [B,3,H,W]); gradients flow to depth (same shape as given) and T.
    Reference: movedepth/trainer.py:519-529 + 535-550, layers.py:646-677."""
    shape = depth.shape
    loss, warped = _Photometric.apply(depth.reshape(shape[0], *shape[-2:]), src, tgt, K, invK, T, ssim_w)
    return loss, warped


def photometric_identity(src, tgt, ssim_w=0.85):
    """SSIM/L1 error between two images without any warp (the auto-masking identity term,
    movedepth/trainer.py:689-693).  No gradient (both inputs are images)."""
    B, _, H, W = src.shape
    src, tgt = _f32(src.detach()).contiguous(), _f32(tgt.detach()).contiguous()
    loss = torch.empty((B, 1, H, W), device=src.device, dtype=torch.float32)
    rc = _lib.lib().mvd_photometric_fwd(_p(None), _p(src), _p(tgt), _p(None), _p(None), _p(None), _p(None), _p(loss), B, H, W,
                                        float(ssim_w), 1, _stream())
    _lib.check(rc, "mvd_photometric_fwd(identity)")
    launch_counter["n"] += 1
    return loss


# ------------------------------------------------------------------------------------- loss glue
class _DispToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, H, W, min_depth, max_depth):
        B, _, hs, ws = disp.shape
        disp = _f32(disp).contiguous()
        inv_far, rng = 1.0 / max_depth, 1.0 / min_depth - 1.0 / max_depth
        depth = torch.empty((B, 1, H, W), device=disp.device, dtype=torch.float32)
        rc = _lib.lib().mvd_disp_to_depth_fwd(_p(disp), _p(depth), B, hs, ws, H, W, inv_far, rng, _stream())
        _lib.check(rc, "mvd_disp_to_depth_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(depth)
        ctx.meta = (B, hs, ws, H, W, rng)
        return depth

    @staticmethod
    def backward(ctx, gdepth):
        depth, = ctx.saved_tensors
        B, hs, ws, H, W, rng = ctx.meta
        gdepth = _f32(gdepth).contiguous()
        gdisp = torch.empty((B, 1, hs, ws), device=depth.device, dtype=torch.float32)
        rc = _lib.lib().mvd_disp_to_depth_bwd(_p(gdepth), _p(depth), _p(gdisp), B, hs, ws, H, W, rng, _stream())
        _lib.check(rc, "mvd_disp_to_depth_bwd")
        launch_counter["n"] += 1
        return gdisp, None, None, None, None


def disp_to_depth_full(disp, height, width, min_depth, max_depth):
    """Sigmoid disparity [B,1,hs,ws] -> depth [B,1,height,width]: bilinear upsampling (align_corners=False) and
    disp_to_depth in one kernel.  Reference: movedepth/trainer.py:512-515, layers.py:400-409."""
    return _DispToDepth.apply(disp, int(height), int(width), float(min_depth), float(max_depth))


class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img, normalize):
        B, _, h, w = disp.shape
        disp, img = _f32(disp).contiguous(), _f32(img.detach()).contiguous()
        assert img.shape == (B, 3, h, w), (disp.shape, img.shape)
        work = torch.empty(B + 2, device=disp.device, dtype=torch.float64)
        loss = torch.empty((), device=disp.device, dtype=torch.float32)
        rc = _lib.lib().mvd_smooth_loss_fwd(_p(disp), _p(img), _p(work), _p(loss), B, h, w, int(normalize), _stream())
        _lib.check(rc, "mvd_smooth_loss_fwd")
        launch_counter["n"] += 3 if normalize else 2
        ctx.save_for_backward(disp, img, work)
        ctx.normalize = int(normalize)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        disp, img, work = ctx.saved_tensors
        B, _, h, w = disp.shape
        gloss = _f32(gloss).contiguous()
        dot = torch.empty(B, device=disp.device, dtype=torch.float64)
        gdisp = torch.empty_like(disp)
        rc = _lib.lib().mvd_smooth_loss_bwd(_p(gloss), _p(disp), _p(img), _p(work), _p(dot), _p(gdisp), B, h, w, ctx.normalize,
                                            _stream())
        _lib.check(rc, "mvd_smooth_loss_bwd")
        launch_counter["n"] += 2
        return gdisp, None, None


def smooth_loss(disp, img, normalize=False):
    """Edge-aware first-order smoothness of `disp` [B,1,h,w] against `img` [B,3,h,w] (movedepth/layers.py:630-643); with
    normalize=True the disparity is first divided by its per-item mean + 1e-7 (movedepth/trainer.py:712-713)."""
    return _SmoothLoss.apply(disp, img, bool(normalize))


class _MaskedSmoothL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, box_xy, H, W, fh, fw, weight):
        B, h, w = a.shape
        a, b = _f32(a).contiguous(), _f32(b).contiguous()
        assert box_xy.dtype == torch.int64 and box_xy.numel() == 2
        sel = torch.empty(B * h * w, device=a.device, dtype=torch.uint8)
        sums = torch.empty(2, device=a.device, dtype=torch.float64)
        loss = torch.empty((), device=a.device, dtype=torch.float32)
        rc = _lib.lib().mvd_masked_smooth_l1_fwd(_p(a), _p(b), _p(box_xy), _p(sel), _p(sums), _p(loss), B, h, w, H, W, fh, fw,
                                                 float(weight), _stream())
        _lib.check(rc, "mvd_masked_smooth_l1_fwd")
        launch_counter["n"] += 2
        ctx.save_for_backward(a, b, sel, sums)
        ctx.weight = float(weight)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        a, b, sel, sums = ctx.saved_tensors
        gloss = _f32(gloss).contiguous()
        ga = torch.empty_like(a)
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        rc = _lib.lib().mvd_masked_smooth_l1_bwd(_p(gloss), _p(a), _p(b), _p(sel), _p(sums), _p(ga), _p(gb), a.numel(), ctx.weight,
                                                 _stream())
        _lib.check(rc, "mvd_masked_smooth_l1_bwd")
        launch_counter["n"] += 1
        return ga, gb, None, None, None, None, None, None


def masked_smooth_l1(a, b, box_xy, height, width, box_h, box_w, weight=1.0):
    """weight * mean over the selected pixels of smooth_l1(a - b); a, b: [B,h,w] depth maps; a low-resolution pixel is
    selected when the bilinear (align_corners=True) resize of the [height,width] box mask (zeros in the box_h x box_w box
    at box_xy = device int64 (x, y)) is non-zero there.  Reference: movedepth/trainer.py:398-400."""
    return _MaskedSmoothL1.apply(a, b, box_xy, int(height), int(width), int(box_h), int(box_w), float(weight))


# ------------------------------------------------------------------------------------- DepthDecoder glue
class _DecoderPrep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, bias, skip, act, up, want_split):
        B, C1, h, w = z.shape
        z = _nhwc(z)
        skip_c = _nhwc(skip) if skip is not None else None
        C2 = skip_c.shape[1] if skip_c is not None else 0
        H, W, C = h * up, w * up, C1 + C2
        if skip_c is not None:
            assert skip_c.shape == (B, C2, H, W), (z.shape, skip.shape, up)
        bias_c = _f32(bias).contiguous() if bias is not None else None
        xp = torch.empty((B, C, H + 2, W + 2), device=z.device, dtype=torch.float32, memory_format=torch.channels_last)
        x3 = torch.empty((B, 3 * C, H + 2, W + 2), device=z.device, dtype=torch.float32,
                         memory_format=torch.channels_last) if want_split else None
        rc = _lib.lib().mvd_decoder_prep_fwd(_p(z), _p(bias_c), _p(skip_c), _p(xp), _p(x3), B, h, w, C1, C2, up, int(act), _stream())
        _lib.check(rc, "mvd_decoder_prep_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(z, bias_c)
        ctx.meta = (B, h, w, C1, C2, up, int(act), skip is not None, bias is not None)
        if x3 is None:
            x3 = z.new_empty(0)
        ctx.mark_non_differentiable(x3)
        return xp, x3

    @staticmethod
    def backward(ctx, gxp, _gx3):
        z, bias_c = ctx.saved_tensors
        B, h, w, C1, C2, up, act, has_skip, has_bias = ctx.meta
        gxp = _nhwc(gxp)
        gz = torch.empty_like(z)                                  # channels-last
        gskip = torch.empty((B, C2, h * up, w * up), device=z.device, dtype=torch.float32,
                            memory_format=torch.channels_last) if has_skip else None
        gbias = torch.empty(C1, device=z.device, dtype=torch.float32) if has_bias else None
        rc = _lib.lib().mvd_decoder_prep_bwd(_p(gxp), _p(z), _p(bias_c), _p(gz), _p(gskip), _p(gbias), B, h, w, C1, C2, up, act,
                                             _stream())
        _lib.check(rc, "mvd_decoder_prep_bwd")
        launch_counter["n"] += 2 if has_skip else 1
        return gz, gbias, gskip, None, None, None


def decoder_prep(z, bias=None, skip=None, act=True, up=1, want_split=False):
    """ReflectionPad2d(1)(cat(nearest_up(act(z + bias), up), skip)) in one kernel, plus (want_split) the 3xTF32 operand
    split of the result.  z: [B,C1,h,w] raw conv output (channels-last storage), skip: [B,C2,up*h,up*w] or None.
    Returns (xp [B,C1+C2,up*h+2,up*w+2], x3 [B,3(C1+C2),...] or an empty tensor).
    Reference: movedepth/networks/depth_decoder.py:72-101, layers.py:521-553, 624-627."""
    return _DecoderPrep.apply(z, bias, skip, bool(act), int(up), bool(want_split))


# ------------------------------------------------------------------------------------- skinny 2-D convolutions (FPN4 / UncertNet)
def conv2d_small_supported(cin, cout, k, stride):
    return bool(_lib.lib().mvd_conv2d_small_supported(int(cin), int(cout), int(k), int(stride)))


class _Conv2dSmall(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, k, stride, pad):
        B, cin, H, W = x.shape
        cout = weight.shape[0]
        x = _nhwc(x)
        w = _f32(weight).contiguous(memory_format=torch.channels_last)           # [cout][k][k][cin] in memory
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty((B, cout, Ho, Wo), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        rc = _lib.lib().mvd_conv2d_small_fwd(_p(x), _p(w), _p(y), B, H, W, cin, cout, k, stride, pad, _stream())
        _lib.check(rc, "mvd_conv2d_small_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(x, w)
        ctx.meta = (B, H, W, cin, cout, k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        B, H, W, cin, cout, k, stride, pad = ctx.meta
        gy = _nhwc(gy)
        gx = gw = None
        L = _lib.lib()
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            _lib.check(L.mvd_conv2d_small_dgrad(_p(gy), _p(w), _p(gx), B, H, W, cin, cout, k, stride, pad, _stream()), "mvd_conv2d_small_dgrad")
            launch_counter["n"] += 1
        if ctx.needs_input_grad[1]:
            nbytes = L.mvd_conv2d_small_wgrad_workspace_bytes(B, H, W, cin, cout, k, stride, pad)
            ws = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
            gw = torch.empty_like(w)                                             # channels-last strides, logical [cout,cin,k,k]
            _lib.check(L.mvd_conv2d_small_wgrad(_p(x), _p(gy), _p(gw), _p(ws), nbytes, B, H, W, cin, cout, k, stride, pad, _stream()),
                       "mvd_conv2d_small_wgrad")
            launch_counter["n"] += 2
        return gx, gw, None, None, None


def conv2d_small(x, weight, k, stride, pad=None):
    """Exact-fp32 direct convolution (no bias) for the skinny FPN4 / UncertNet layers (zero padding k//2) and the DepthDecoder's
    finest stage (pad=0: pre-padded input); x [B,cin,H,W] (channels-last storage), weight [cout,cin,k,k].
    Reference: movedepth/networks/resnet_encoder.py:325-341, 453-475; depth_decoder.py:72-101."""
    return _Conv2dSmall.apply(x, weight, int(k), int(stride), int(k) // 2 if pad is None else int(pad))


# ------------------------------------------------------------------------------------- pose-net outputs -> camera transform
class _PoseMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        B = axisangle.shape[0]
        aa = _f32(axisangle).reshape(B, 3).contiguous()
        tr = _f32(translation).reshape(B, 3).contiguous()
        M = torch.empty((B, 4, 4), device=aa.device, dtype=torch.float32)
        _lib.check(_lib.lib().mvd_pose_matrix_fwd(_p(aa), _p(tr), _p(M), B, int(invert), _stream()), "mvd_pose_matrix_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(aa, tr)
        ctx.meta = (bool(invert), tuple(axisangle.shape), tuple(translation.shape))
        return M

    @staticmethod
    def backward(ctx, gM):
        aa, tr = ctx.saved_tensors
        invert, sa, st = ctx.meta
        B = aa.shape[0]
        gaa, gtr = torch.empty_like(aa), torch.empty_like(tr)
        _lib.check(_lib.lib().mvd_pose_matrix_bwd(_p(aa), _p(tr), _p(_f32(gM).contiguous()), _p(gaa), _p(gtr), B, int(invert), _stream()),
                   "mvd_pose_matrix_bwd")
        launch_counter["n"] += 1
        return gaa.reshape(sa), gtr.reshape(st), None


def pose_matrix(axisangle, translation, invert=False):
    """transformation_from_parameters (movedepth/layers.py:412-429) in one launch per direction: axis-angle [B,1,3] / [B,3] and
    translation -> [B,4,4] camera transform (T(t) R, or R^T T(-t) with invert)."""
    return _PoseMatrix.apply(axisangle, translation, bool(invert))


# ------------------------------------------------------------------------------------- ResNet stem max-pool
class _MaxPool3x3S2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, C, H, W = x.shape
        x = _nhwc(x)
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        idx = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.uint8)
        rc = _lib.lib().mvd_maxpool3x3s2_fwd(_p(x), _p(y), _p(idx), B, H, W, C, _stream())
        _lib.check(rc, "mvd_maxpool3x3s2_fwd")
        launch_counter["n"] += 1
        ctx.save_for_backward(idx)
        ctx.meta = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        idx, = ctx.saved_tensors
        B, C, H, W = ctx.meta
        gy = _nhwc(gy)
        gx = torch.empty((B, C, H, W), device=gy.device, dtype=torch.float32, memory_format=torch.channels_last)
        rc = _lib.lib().mvd_maxpool3x3s2_bwd(_p(gy), _p(idx), _p(gx), B, H, W, C, _stream())
        _lib.check(rc, "mvd_maxpool3x3s2_bwd")
        launch_counter["n"] += 1
        return gx


def maxpool3x3s2(x):
    """MaxPool2d(3, stride 2, padding 1) of the ResNet stem on a channels-last activation (C % 4 == 0)."""
    return _MaxPool3x3S2.apply(x)
