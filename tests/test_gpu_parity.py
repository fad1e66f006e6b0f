"""GPU parity, kernel level: every hand-written kernel, called through the C ABI (ctypes), against the CPU oracle
on the same seeded inputs and against the reference-generated golden vectors.  Tolerances are stated per test.
The whole-step / whole-model tests live in test_gpu_step.py, which is collected AFTER this file, so a step-level
failure cannot hide a kernel test."""
import os

import numpy as np
import pytest
import torch

import _cases as C
from _weights import fill_deterministic
from oracle import layers as OL

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from movedepth_b200 import ops as _ops
    _ops._lib.lib()
    return _ops


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLD, "ops.npz")))


def g(t):
    return t.to(DEV)


def grouped_oracle(c, G=16):
    ref, src = c["ref"].clone().requires_grad_(True), c["src"].clone().requires_grad_(True)
    vol = OL.group_correlation(OL.cost_volume(ref, src, c["K"], c["invK"], c["hyps"], c["pose"]), G)   # [B,D,G,h,w]
    return ref, src, vol


# ---------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("name", C.COSTVOL_CASES)
@pytest.mark.parametrize("flags", [0, 1, 2, 3], ids=["tma", "gather", "tma-notable", "gather-notable"])
@pytest.mark.parametrize("layout", [0, 1], ids=["bgdhw", "bdhwg"])
def test_costvol_grouped_forward_backward(ops, name, flags, layout):
    """fused warp+gather+group-correlation vs oracle generate_costvol + group mean.
    atol 2e-4: sampling coordinates differ by <=6e-5 px from the reference's normalise/unnormalise
    round trip on smooth N(0,1) features."""
    c = C.case_costvol(name)
    ref_o, src_o, want = grouped_oracle(c)
    gv = torch.randn(want.shape, generator=torch.Generator().manual_seed(5))
    (want * gv).sum().backward()
    ref, src = g(c["ref"]).requires_grad_(True), g(c["src"]).requires_grad_(True)
    got = ops.costvol_grouped(ref, src, g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]), prior=g(c["prior"]),
                              ratio=g(c["ratio"]), layout=layout, flags=flags)
    assert got.shape == (c["B"], 16, c["D"], c["h"], c["w"])
    torch.testing.assert_close(got.permute(0, 2, 1, 3, 4).cpu(), want.detach(), atol=2e-4, rtol=1e-4)
    (got * g(gv).permute(0, 2, 1, 3, 4)).sum().backward()
    scale = float(ref_o.grad.abs().max())
    torch.testing.assert_close(ref.grad.cpu(), ref_o.grad, atol=2e-4 * scale, rtol=1e-3)
    torch.testing.assert_close(src.grad.cpu(), src_o.grad, atol=2e-4 * float(src_o.grad.abs().max()), rtol=1e-3)


def test_costvol_grouped_explicit_hypotheses_equal_ratio_form(ops):
    c = C.case_costvol("forward")
    args = (g(c["ref"]), g(c["src"]), g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]))
    a = ops.costvol_grouped(*args, prior=g(c["prior"]), ratio=g(c["ratio"]))
    b = ops.costvol_grouped(*args, hyps=g(c["hyps"]))
    torch.testing.assert_close(a, b, atol=1e-6, rtol=1e-6)


def test_costvol_grouped_tma_and_gather_routes_agree_bitwise(ops):
    c = C.case_costvol("sideways", B=3, h=24, w=96, D=24)
    args = (g(c["ref"]), g(c["src"]), g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]))
    outs = [ops.costvol_grouped(*args, prior=g(c["prior"]), ratio=g(c["ratio"]), flags=f, layout=l)
            for f in (0, 1, 2, 3) for l in (0, 1)]          # TMA / gather x table / direct x both layouts
    for o in outs[1:]:
        assert torch.equal(outs[0], o)


def test_costvol_grouped_more_batch_items_than_one_launch_holds(ops):
    """The forward kernel keeps the geometry of <= 16 batch items in shared memory; larger batches are split."""
    c = C.case_costvol("forward", B=19, h=8, w=64, D=8)
    _, _, want = grouped_oracle(c)
    for flags in (0, 16):                                   # TMA stores / plain stores
        got = ops.costvol_grouped(g(c["ref"]), g(c["src"]), g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]),
                                  prior=g(c["prior"]), ratio=g(c["ratio"]), layout=1, flags=flags)
        torch.testing.assert_close(got.permute(0, 2, 1, 3, 4).cpu(), want.detach(), atol=2e-4, rtol=1e-4)


def test_costvol_grouped_identity_pose_is_plain_correlation(ops):
    """SURVEY section 4 known answer: identity pose -> ref*src broadcast over D."""
    c = C.case_costvol("identity", B=1, h=16, w=64, D=8)
    got = ops.costvol_grouped(g(c["ref"]), g(c["src"]), g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]),
                              prior=g(c["prior"]), ratio=g(c["ratio"]))
    prod = (c["ref"] * c["src"]).view(1, 2, 16, 16, 64).mean(1)                      # group g = {g, g+16}
    torch.testing.assert_close(got.cpu(), prod[:, :, None].expand(-1, -1, 8, -1, -1), atol=1e-3, rtol=0)


def test_costvol_grouped_ragged_shapes(ops):
    """width not a multiple of the 32-pixel tile, odd height, D not a multiple of the chunk count."""
    c = C.case_costvol("forward", B=1, h=7, w=45, D=5)
    _, _, want = grouped_oracle(c)
    for flags in (0, 1, 2):
        for layout in (0, 1):
            got = ops.costvol_grouped(g(c["ref"]), g(c["src"]), g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]),
                                      prior=g(c["prior"]), ratio=g(c["ratio"]), flags=flags, layout=layout)
            torch.testing.assert_close(got.permute(0, 2, 1, 3, 4).cpu(), want.detach(), atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("name,hw,D", [("forward", (48, 160), 96), ("stress", (48, 160), 96), ("sideways", (80, 256), 128)])
def test_costvol_grouped_matches_oracle_at_benchmark_shapes(ops, name, hw, D):
    """One frame at the BASELINE config-2 / config-5 feature shapes (48x160 D=96, 80x256 D=128) against the oracle,
    both layouts; `stress` is the degenerate pose whose footprints overflow the table and the TMA box."""
    c = C.case_costvol(name, B=1, h=hw[0], w=hw[1], D=D)
    with torch.no_grad():
        want = OL.group_correlation(OL.cost_volume(c["ref"], c["src"], c["K"], c["invK"], c["hyps"], c["pose"]), 16)
    scale = float(want.abs().max())
    for layout in (0, 1):
        got = ops.costvol_grouped(g(c["ref"]), g(c["src"]), g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]),
                                  prior=g(c["prior"]), ratio=g(c["ratio"]), layout=layout)
        torch.testing.assert_close(got.permute(0, 2, 1, 3, 4).cpu(), want, atol=2e-4 * max(1.0, scale), rtol=1e-4)


def test_costvol_grouped_linearity_at_full_size(ops):
    """Size-independent property at BASELINE config 2 (B=6, 48x160, D=96): the volume is bilinear in
    (ref, src): V(a*ref, src1+src2) == a*(V(ref,src1) + V(ref,src2))."""
    c = C.case_costvol("forward", B=6, h=48, w=160, D=96)
    geo = (g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]))
    kw = dict(prior=g(c["prior"]), ratio=g(c["ratio"]))
    ref, s1 = g(c["ref"]), g(c["src"])
    s2 = g(C.smooth_noise((6, 32, 48, 160), 99))
    lhs = ops.costvol_grouped(2.0 * ref, s1 + s2, *geo, **kw)
    rhs = 2.0 * (ops.costvol_grouped(ref, s1, *geo, **kw) + ops.costvol_grouped(ref, s2, *geo, **kw))
    torch.testing.assert_close(lhs, rhs, atol=1e-4, rtol=1e-4)
    assert lhs.shape == (6, 16, 96, 48, 160)


@pytest.mark.parametrize("name,shape", [("forward", (2, 7, 45, 5)), ("sideways", (19, 8, 64, 8)), ("stress", (1, 48, 160, 96)),
                                        ("forward", (2, 48, 160, 96))], ids=["ragged", "19-items", "stress-full", "config2"])
def test_costvol_backward_streaming_kernel_agrees_with_the_round1_kernel(ops, name, shape):
    """K1b v5 (warp-autonomous: lane pair per pixel, long hypothesis chunks, gradients streamed through a per-warp bulk-copy ring
    in the channels-last layout / four hypotheses in flight through registers otherwise) vs the round-1 kernel (MVD_FLAG_BWD_V2:
    eight short chunks per pixel, one hypothesis in flight) on the same inputs: same sums in a different order.  Repeated: the
    ring's refill race showed up in ~7 % of runs before the loads were forced to complete ahead of the refill."""
    B, h, w, D = shape
    c = C.case_costvol(name, B=B, h=h, w=w, D=D)
    gv = g(torch.randn((B, 16, D, h, w), generator=torch.Generator().manual_seed(6)))
    grads = []
    for flags in (32, 0, 0x100, 0x400, 0, 0, 0, 0):                    # v2 / v5 (3 chunks) / 1 chunk / 4 chunks / v5 again x4
        for layout in (0, 1):
            ref, src = g(c["ref"]).requires_grad_(True), g(c["src"]).requires_grad_(True)
            got = ops.costvol_grouped(ref, src, g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]), prior=g(c["prior"]), ratio=g(c["ratio"]),
                                      layout=layout, flags=flags)
            (got * gv).sum().backward()
            grads.append((ref.grad, src.grad))
    for gr, gs in grads[1:]:
        torch.testing.assert_close(gr, grads[0][0], atol=2e-5 * float(grads[0][0].abs().max()), rtol=1e-4)
        torch.testing.assert_close(gs, grads[0][1], atol=2e-5 * float(grads[0][1].abs().max()), rtol=1e-4)


def test_costvol_backward_euler_identity_at_full_size(ops):
    """Size-independent property of the adjoint at BASELINE config 2 (B=6, 48x160, D=96): the volume is linear in ref and in
    src, so <d ref, ref> = <d src, src> = <V, G> for any upstream gradient G (fp64 sums of fp32 products)."""
    c = C.case_costvol("forward", B=6, h=48, w=160, D=96)
    ref, src = g(c["ref"]).requires_grad_(True), g(c["src"]).requires_grad_(True)
    vol = ops.costvol_grouped(ref, src, g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]), prior=g(c["prior"]), ratio=g(c["ratio"]), layout=1)
    gv = torch.randn(vol.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(8))
    (vol * gv).sum().backward()
    want = float((vol.detach().double() * gv.double()).sum())
    a = float((ref.grad.double() * ref.detach().double()).sum())
    b = float((src.grad.double() * src.detach().double()).sum())
    scale = float((vol.detach().double() * gv.double()).abs().sum())
    assert abs(a - want) < 1e-6 * scale and abs(b - want) < 1e-6 * scale, (a, b, want, scale)


@pytest.mark.parametrize("name", C.COSTVOL_CASES)
def test_costvol_full_matches_reference_golden(ops, gold, name):
    """public generate_costvol layout [B,D,C,h,w] against the reference's own output."""
    from movedepth_b200 import layers as PL
    c = C.case_costvol(name)
    ref, src = g(c["ref"]).requires_grad_(True), g(c["src"]).requires_grad_(True)
    vol = PL.generate_costvol(ref, src, g(c["K"]), g(c["invK"]), g(c["hyps"]), g(c["pose"]), c["D"], None, None)
    np.testing.assert_allclose(vol.detach().cpu().numpy(), gold["costvol_%s" % name], atol=2e-4, rtol=1e-4)
    (vol * g(c["gvol"])).sum().backward()
    for got, key in ((ref.grad, "gref"), (src.grad, "gsrc")):
        want = gold["costvol_%s_%s" % (name, key)]
        np.testing.assert_allclose(got.cpu().numpy(), want, atol=3e-4 * np.abs(want).max(), rtol=1e-3)


# ---------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("radius", [1, 2])
def test_regress_matches_oracle_and_golden(ops, gold, radius):
    c = C.case_localmax()
    logits = torch.log(c["prob"])
    lg = g(logits).requires_grad_(True)
    prob, ent, depth = ops.regress_depth(lg, g(c["inv_a"]), g(c["inv_b"]), radius, want_prob=True)
    np.testing.assert_allclose(prob.cpu().numpy(), c["prob"].numpy(), atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(ent.detach().cpu().numpy(), gold["entropy"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), gold["localmax_r%d" % radius], atol=0, rtol=1e-5)
    # backward vs autograd through the oracle
    lo = logits.clone().requires_grad_(True)
    p = torch.softmax(lo, 1)
    gen = torch.Generator().manual_seed(3)
    ge, gd = torch.randn(ent.shape, generator=gen), torch.randn(depth.shape, generator=gen)
    ((OL.entropy(p, 1, True) * ge).sum() + (OL.localmax(p, radius, c["D"], c["inv_a"], c["inv_b"]) * gd).sum()).backward()
    ((ent * g(ge)).sum() + (depth * g(gd)).sum()).backward()
    torch.testing.assert_close(lg.grad.cpu(), lo.grad, atol=1e-5 * float(lo.grad.abs().max()), rtol=1e-3)


def test_localmax_public_signature_and_known_answer(ops, gold):
    from movedepth_b200 import layers as PL
    c = C.case_localmax()
    got = PL.localmax(g(c["onehot"]), 1, c["D"], g(c["inv_a"]), g(c["inv_b"]))
    np.testing.assert_allclose(got.cpu().numpy(), gold["localmax_onehot"], atol=0, rtol=1e-5)
    # one-hot at i picks hypothesis D-1-i
    D = 16
    hyp = OL.depth_hypotheses(torch.full((1, 1, 1, 1), 10.0), D, 0.3)
    for i, want in ((0, 10 / 1.3), (5, 8.9041), (15, 13.0)):
        p = torch.zeros(1, D, 1, 1)
        p[0, i] = 1
        d = PL.localmax(g(p), 1, D, g(1 / hyp[:, -1]), g(1 / hyp[:, 0]))
        assert abs(float(d) - want) < 1e-3


def test_regress_window_clamps_at_both_ends(ops):
    """argmax at index 0 / D-1: clamped duplicates are counted twice (reference behaviour)."""
    D = 6
    for top in (0, D - 1):
        logits = torch.randn(1, D, 2, 3, generator=torch.Generator().manual_seed(top))
        logits[:, top] += 6.0
        inv_a, inv_b = torch.full((1, 2, 3), 0.1), torch.full((1, 2, 3), 0.4)
        lg = g(logits).requires_grad_(True)
        _, _, depth = ops.regress_depth(lg, g(inv_a), g(inv_b), 2)
        lo = logits.clone().requires_grad_(True)
        want = OL.localmax(torch.softmax(lo, 1), 2, D, inv_a, inv_b)
        torch.testing.assert_close(depth.cpu(), want.detach(), atol=0, rtol=1e-5)
        want.sum().backward()
        depth.sum().backward()
        torch.testing.assert_close(lg.grad.cpu(), lo.grad, atol=1e-6, rtol=1e-3)


# ---------------------------------------------------------------------------------------------- K4
def test_convex_upsample_matches_golden_and_autograd(ops, gold):
    c = C.case_convex()
    d, m = g(c["depth"]).requires_grad_(True), g(c["mask"]).requires_grad_(True)
    up = ops.convex_upsample(d, m, 2)
    np.testing.assert_allclose(up.detach().cpu().numpy(), gold["convex_up"], atol=1e-5, rtol=1e-5)
    do, mo = c["depth"].clone().requires_grad_(True), c["mask"].clone().requires_grad_(True)
    gu = torch.randn(up.shape, generator=torch.Generator().manual_seed(2))
    (OL.convex_upsample(do, mo, 2) * gu).sum().backward()
    (up * g(gu)).sum().backward()
    torch.testing.assert_close(d.grad.cpu(), do.grad, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(m.grad.cpu(), mo.grad, atol=1e-5, rtol=1e-4)


# ---------------------------------------------------------------------------------------------- Adam
def test_fused_adam_matches_torch_adam(ops):
    gen = torch.Generator().manual_seed(7)
    n = 10007                                           # not a multiple of 4: exercises the tail
    p0 = torch.randn(n, generator=gen)
    want = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([want], lr=2e-4)
    p, m, v = g(p0.clone()), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        grad = torch.randn(n, generator=gen)
        want.grad = grad.clone()
        opt.step()
        ops.adam_step(p, g(grad * 2), m, v, step, 2e-4, grad_scale=0.5)
        torch.testing.assert_close(p.cpu(), want.detach(), atol=1e-7, rtol=1e-6)


# ---------------------------------------------------------------------------------------------- reg3d output head
@pytest.mark.parametrize("shape", [(2, 8, 8, 32), (1, 5, 11, 45), (2, 24, 24, 80)], ids=["aligned", "ragged", "multi-tile"])
def test_prob_conv3d_matches_torch_conv3d(ops, shape):
    """Conv3d(16->1, 3x3x3, pad 1): hand-written stencil kernels (fwd, dgrad, wgrad) vs torch's CPU conv3d in fp64.
    fp32 accumulation of 432 products: 1e-5 relative to the output scale."""
    import torch.nn.functional as F
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(B, 16, D, H, W, generator=gen)
    w = torch.randn(1, 16, 3, 3, 3, generator=gen) * 0.1
    gy = torch.randn(B, 1, D, H, W, generator=gen)
    xo, wo = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yo = F.conv3d(xo, wo, padding=1)
    (yo * gy.double()).sum().backward()
    xg = g(x).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    wg = g(w).requires_grad_(True)
    y = ops.conv3d_c16_to_1(xg, wg)
    assert y.shape == (B, 1, D, H, W)
    (y * g(gy)).sum().backward()
    for got, want in ((y, yo), (xg.grad, xo.grad), (wg.grad, wo.grad)):
        want = want.detach().float()
        torch.testing.assert_close(got.detach().cpu(), want, atol=1e-5 * float(want.abs().max()), rtol=1e-5)


def test_reg3d_uses_the_stencil_head_and_matches_the_oracle(ops):
    """movedepth_b200 reg3d (fp32 policy) vs the oracle's reg3d on the same weights: logits and input gradient."""
    from movedepth_b200 import networks as PN, precision as PR
    from oracle import networks as ON
    PR.set_policy("fp32")
    gen = torch.Generator().manual_seed(5)
    vol = torch.randn(1, 16, 8, 16, 32, generator=gen)
    a, b = PN.reg3d(16, 16, 3), ON.Reg3d(16, 16, 3)
    fill_deterministic(a)
    fill_deterministic(b)
    a.to(DEV)
    xa = g(vol).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    xb = vol.clone().requires_grad_(True)
    ya, yb = a.forward_volume(xa), b(xb.permute(0, 2, 1, 3, 4))
    torch.testing.assert_close(ya.detach().cpu(), yb.detach(), atol=1e-4 * float(yb.abs().max()), rtol=1e-4)
    gy = torch.randn(yb.shape, generator=gen)
    (ya * g(gy)).sum().backward()
    (yb * gy).sum().backward()
    torch.testing.assert_close(xa.grad.cpu(), xb.grad, atol=1e-3 * float(xb.grad.abs().max()), rtol=1e-3)
    torch.testing.assert_close(a.prob.weight.grad.cpu(), b.prob.weight.grad, atol=1e-3 * float(b.prob.weight.grad.abs().max()), rtol=1e-3)


@pytest.mark.parametrize("shape", [(1, 6, 8, 32), (1, 5, 11, 45), (2, 26, 24, 80)], ids=["aligned", "ragged", "multi-tile"])
def test_conv3d_16_to_16_matches_torch_conv3d(ops, shape):
    """Conv3d(16->16, 3x3x3, pad 1) tensor-core implicit GEMM vs torch's CPU conv3d in fp64: the 3xTF32 forward within
    2e-5 of the output scale (fp32-class), the single-pass TF32 forward and data gradient within 2e-3 (TF32-class)."""
    import torch.nn.functional as F
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(B, 16, D, H, W, generator=gen)
    w = torch.randn(16, 16, 3, 3, 3, generator=gen) * 0.1
    gy = torch.randn(B, 16, D, H, W, generator=gen)
    xo = x.double().requires_grad_(True)
    yo = F.conv3d(xo, w.double(), padding=1)
    (yo * gy.double()).sum().backward()
    yo = yo.detach().float()
    xg = g(x).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    wg = g(w).requires_grad_(True)
    y3 = ops.conv3d_c16_to_16(xg, wg, 3)
    torch.testing.assert_close(y3.detach().cpu(), yo, atol=2e-5 * float(yo.abs().max()), rtol=0)
    y1 = ops.conv3d_c16_to_16(xg, wg, 1)
    torch.testing.assert_close(y1.detach().cpu(), yo, atol=3e-3 * float(yo.abs().max()), rtol=0)
    ym = ops.c16c16_conv(xg, wg, 0, 3)                                     # the mma.sync variant of the same contract
    torch.testing.assert_close(ym.cpu(), yo, atol=2e-5 * float(yo.abs().max()), rtol=0)
    (y3 * g(gy)).sum().backward()
    gxo = xo.grad.float()
    torch.testing.assert_close(xg.grad.cpu(), gxo, atol=3e-3 * float(gxo.abs().max()), rtol=0)
    gwo = None
    wo = w.double().requires_grad_(True)
    (F.conv3d(x.double(), wo, padding=1) * gy.double()).sum().backward()
    gwo = wo.grad.float()
    torch.testing.assert_close(wg.grad.cpu(), gwo, atol=3e-3 * float(gwo.abs().max()), rtol=0)     # TF32 weight gradient (tcgen05; odd W: fp32)


# ---------------------------------------------------------------------------------------------- BatchNorm kernels
@pytest.mark.parametrize("shape,relu,res", [((3, 16, 5, 6, 7), True, False), ((4, 64, 9, 11), True, True), ((2, 8, 6, 10), False, False),
                                            ((2, 512, 3, 5), True, True), ((2, 32, 4, 6, 8), True, "post")],
                         ids=["3d-relu", "2d-relu-res", "2d-plain", "c512", "3d-relu-then-skip"])
@pytest.mark.parametrize("one_kernel", [False, True], ids=["three-kernels", "one-kernel"])
def test_fused_batchnorm_matches_torch(shape, relu, res, one_kernel, monkeypatch):
    """csrc/bn.cu (stats / finalize / apply, backward reduce / apply; and the one-kernel forms with the grid barrier) vs torch's
    batch_norm + add + relu in fp64: outputs, input / residual / weight / bias gradients and the running statistics."""
    import torch.nn as nn
    from movedepth_b200 import norm as NM
    monkeypatch.setattr(NM, "fuse_bytes", (1 << 30) if one_kernel else 0)
    gen = torch.Generator().manual_seed(9)
    C = shape[1]
    cls = nn.BatchNorm3d if len(shape) == 5 else nn.BatchNorm2d
    fmt = torch.channels_last_3d if len(shape) == 5 else torch.channels_last
    x = torch.randn(shape, generator=gen) * 2 + 0.5
    r = torch.randn(shape, generator=gen) if res else None
    gy = torch.randn(shape, generator=gen)
    ref = cls(C).double()
    mine = cls(C).to(DEV)
    with torch.no_grad():
        ref.weight.copy_(1 + 0.2 * torch.randn(C, generator=gen))
        ref.bias.copy_(0.1 * torch.randn(C, generator=gen))
        mine.weight.copy_(ref.weight.float())
        mine.bias.copy_(ref.bias.float())
    xo = x.double().requires_grad_(True)
    ro = r.double().requires_grad_(True) if res else None
    yo = ref(xo)
    if res == "post":                                # U-Net skip: ReLU first, then the addition (reg3d conv7/9/11)
        yo = torch.relu(yo) + ro
    else:
        if res:
            yo = yo + ro
        if relu:
            yo = torch.relu(yo)
    (yo * gy.double()).sum().backward()
    xg = g(x).contiguous(memory_format=fmt).requires_grad_(True)
    rg = g(r).contiguous(memory_format=fmt).requires_grad_(True) if res else None
    y = NM.bn_act(mine, xg, relu=relu, residual=rg, post=(res == "post"))
    (y * g(gy)).sum().backward()

    def close(a, b, tol=2e-5):
        b = b.detach().float()
        torch.testing.assert_close(a.detach().cpu(), b, atol=tol * max(1.0, float(b.abs().max())), rtol=tol)
    close(y, yo)
    close(xg.grad, xo.grad, 1e-4)
    if res:
        close(rg.grad, ro.grad)
    close(mine.weight.grad, ref.weight.grad, 1e-4)
    close(mine.bias.grad, ref.bias.grad, 1e-4)
    close(mine.running_mean, ref.running_mean)
    close(mine.running_var, ref.running_var)
    assert int(mine.num_batches_tracked) == 1
    if one_kernel:                                   # the workspace is handed back zeroed, no timeout was recorded
        torch.cuda.synchronize()
        assert int(NM.workspace(DEV, 0).view(torch.int64).abs().sum()) == 0


def test_one_kernel_batchnorm_equals_the_three_kernel_form_at_a_full_grid():
    """A ResNet layer1-sized activation (the barrier runs at its full grid of 2/3 of the SMs, many rows per thread, both
    unrolled and tail iterations), repeated with different channel counts on one shared workspace: outputs, saved statistics
    and gradients of the one-kernel BatchNorm vs the stats / finalize / apply kernels (fp64 atomics in a different order:
    1e-6), running statistics after three steps, workspace zero afterwards."""
    import torch.nn as nn
    from movedepth_b200 import norm as NM
    gen = torch.Generator(device=DEV).manual_seed(77)
    for shape, relu, res in [((6, 64, 48, 160), True, True), ((3, 8, 95, 161), True, False), ((2, 128, 24, 80), False, False),
                             ((1, 16, 7, 24, 80), True, "post")]:
        C = shape[1]
        fmt = torch.channels_last_3d if len(shape) == 5 else torch.channels_last
        cls = nn.BatchNorm3d if len(shape) == 5 else nn.BatchNorm2d
        bns = [cls(C).to(DEV) for _ in range(2)]
        with torch.no_grad():
            bns[0].weight.copy_(1 + 0.2 * torch.randn(C, device=DEV, generator=gen))
            bns[0].bias.copy_(0.1 * torch.randn(C, device=DEV, generator=gen))
            bns[1].load_state_dict(bns[0].state_dict())
        for it in range(3):
            x = (torch.randn(shape, device=DEV, generator=gen) * 2 + 0.5).contiguous(memory_format=fmt)
            r = torch.randn(shape, device=DEV, generator=gen).contiguous(memory_format=fmt) if res else None
            gy = torch.randn(shape, device=DEV, generator=gen).contiguous(memory_format=fmt)
            outs = []
            for k, limit in enumerate((0, 1 << 30)):
                NM.fuse_bytes, keep = limit, NM.fuse_bytes
                try:
                    xi = x.clone().requires_grad_(True)
                    ri = r.clone().requires_grad_(True) if res else None
                    y = NM.bn_act(bns[k], xi, relu=relu, residual=ri, post=(res == "post"))
                    (y * gy).sum().backward()
                finally:
                    NM.fuse_bytes = keep
                outs.append((y.detach(), xi.grad, ri.grad if res else None, bns[k].weight.grad.clone(), bns[k].bias.grad.clone()))
                bns[k].zero_grad()
            for a, b in zip(*outs):
                if a is not None:
                    torch.testing.assert_close(b, a, atol=1e-5 * max(1.0, float(a.abs().max())), rtol=1e-5)
        for name in ("running_mean", "running_var"):
            torch.testing.assert_close(getattr(bns[1], name), getattr(bns[0], name), atol=1e-6, rtol=1e-6)
        assert int(bns[1].num_batches_tracked) == 3
    torch.cuda.synchronize()
    assert int(NM.workspace(DEV, 0).view(torch.int64).abs().sum()) == 0


def _sync_bn_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.nn as nn
    from movedepth_b200 import norm as NM
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)      # gloo stages the 2C fp64 sums through the host
    torch.cuda.set_device(0)
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(4, 16, 6, 10, generator=gen) * 1.5 + 0.3
    gy = torch.randn(4, 16, 6, 10, generator=gen)
    bn = nn.SyncBatchNorm(16).to("cuda:0")
    xs = x[rank * 2:rank * 2 + 2].to("cuda:0").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = NM.bn_act(bn, xs, relu=True)
    (y * gy[rank * 2:rank * 2 + 2].to("cuda:0")).sum().backward()
    # numpy arrays travel by value: tensors would be shared through file descriptors that die with this process
    q.put((rank, y.detach().cpu().numpy(), xs.grad.cpu().numpy(), bn.weight.grad.cpu().numpy(), bn.running_var.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_batchnorm_statistics_are_exchanged_across_ranks():
    """nn.SyncBatchNorm semantics (movedepth/trainer.py:69-129): two ranks, each with half of the batch, must produce
    what one BatchNorm over the whole batch produces; the weight gradients of the ranks add up to the full one."""
    import torch.multiprocessing as mp
    import torch.nn as nn
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_bn_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, [torch.from_numpy(t) for t in rest]) for r, *rest in [q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
    gen = torch.Generator().manual_seed(21)
    x = (torch.randn(4, 16, 6, 10, generator=gen) * 1.5 + 0.3).double().requires_grad_(True)
    gy = torch.randn(4, 16, 6, 10, generator=gen).double()
    bn = nn.BatchNorm2d(16).double()
    y = torch.relu(bn(x))
    (y * gy).sum().backward()
    for r in range(2):
        yr, gxr, gwr, rvr = got[r]
        torch.testing.assert_close(yr, y[r * 2:r * 2 + 2].detach().float(), atol=2e-5, rtol=2e-5)
        torch.testing.assert_close(gxr, x.grad[r * 2:r * 2 + 2].float(), atol=1e-4, rtol=1e-4)
        torch.testing.assert_close(rvr, bn.running_var.float(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(got[0][2] + got[1][2], bn.weight.grad.float(), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("shape", [(1, 6, 7, 32), (1, 5, 11, 45), (2, 26, 24, 80)], ids=["aligned", "ragged", "multi-tile"])
def test_conv3d_16_to_16_tcgen05_matches_torch_conv3d(ops, shape):
    """The tcgen05 / TMEM version of the 16->16 layer (every tap = the TMA-staged slice at a shifted descriptor start
    address) vs torch's CPU conv3d in fp64: 3xTF32 forward and data gradient within 2e-5 of the output scale, single-pass
    TF32 (operands truncated to 10 mantissa bits by the tensor core) within 3e-3."""
    import torch.nn.functional as F
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(B, 16, D, H, W, generator=gen)
    w = torch.randn(16, 16, 3, 3, 3, generator=gen) * 0.1
    yo = F.conv3d(x.double(), w.double(), padding=1).float()
    go = torch.nn.grad.conv3d_input(x.shape, w.double(), x.double(), padding=1).float()      # data gradient with gy := x
    xg, wg = g(x).contiguous(memory_format=torch.channels_last_3d), g(w)
    for passes, tol in ((3, 2e-5), (1, 3e-3)):
        y = ops.c16c16_conv_tc(xg, wg, 0, passes)
        torch.testing.assert_close(y.cpu(), yo, atol=tol * float(yo.abs().max()), rtol=0)
        gx = ops.c16c16_conv_tc(xg, wg, 1, passes)
        torch.testing.assert_close(gx.cpu(), go, atol=tol * float(go.abs().max()), rtol=0)


@pytest.mark.parametrize("shape", [(1, 6, 7, 32), (1, 5, 11, 46), (2, 26, 24, 80)], ids=["aligned", "ragged", "multi-tile"])
def test_conv3d_16_to_16_tcgen05_weight_gradient(ops, shape):
    """tcgen05 weight gradient (MN-major SWIZZLE_128B_BASE32B operands through a paired TMA view, even + odd position
    halves summed) vs torch's fp64 conv3d_weight: single-pass TF32 with truncated operands -> 3e-3 of the scale; the
    exact-fp32 FFMA2 kernel of the same contract -> 2e-5.  Odd widths fall back to the FFMA2 kernel."""
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(B, 16, D, H, W, generator=gen)
    gy = torch.randn(B, 16, D, H, W, generator=gen)
    want = torch.nn.grad.conv3d_weight(x.double(), (16, 16, 3, 3, 3), gy.double(), padding=1).float()
    xg = g(x).contiguous(memory_format=torch.channels_last_3d)
    gg = g(gy).contiguous(memory_format=torch.channels_last_3d)
    scale = float(want.abs().max())
    torch.testing.assert_close(ops.c16c16_wgrad_tc(gg, xg).cpu(), want, atol=3e-3 * scale, rtol=0)
    torch.testing.assert_close(ops.c16c16_wgrad(gg, xg).cpu(), want, atol=2e-5 * scale, rtol=0)
    xo = g(torch.randn(1, 16, 4, 7, 33, generator=gen)).contiguous(memory_format=torch.channels_last_3d)
    go = g(torch.randn(1, 16, 4, 7, 33, generator=gen)).contiguous(memory_format=torch.channels_last_3d)
    torch.testing.assert_close(ops.c16c16_wgrad_tc(go, xo), ops.c16c16_wgrad(go, xo))


# ---------------------------------------------------------------------------------------------- loss assembly
@pytest.mark.parametrize("nsrc,automask", [(1, True), (2, True), (2, False), (1, False)])
def test_reproj_select_matches_the_tensor_formula(ops, nsrc, automask):
    """min over sources + identity auto-mask + masked mean (trainer.py:687-709) in one kernel vs the tensor code it
    replaces, values and the gradients routed to the selected source."""
    gen = torch.Generator().manual_seed(12)
    B, H, W = 2, 24, 40
    ls = [torch.rand(B, 1, H, W, generator=gen) for _ in range(nsrc)]
    ident = torch.rand(B, 1, H, W, generator=gen) * 0.8
    nz = torch.randn(B, 1, H, W, generator=gen)
    want_in = [l.clone().requires_grad_(True) for l in ls]
    reproj = torch.cat(want_in, 1).min(1, keepdim=True)[0]
    mask = (reproj <= ident + nz * 1e-5).float() if automask else torch.ones_like(reproj)
    want = (reproj * mask).sum() / (mask.sum() + 1e-7)
    (want * 3.0).backward()
    got_in = [g(l).requires_grad_(True) for l in ls]
    loss, rp = ops.reproj_select(got_in, g(ident) if automask else None, g(nz) if automask else None)
    (loss * 3.0).backward()
    torch.testing.assert_close(rp.cpu(), reproj.detach())
    torch.testing.assert_close(loss.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-7)
    for a, b in zip(got_in, want_in):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-9)


def test_reg3d_first_layer_kernels_are_linear_at_full_size(ops):
    """Size-independent property at BASELINE config 2 (volume [6,16,96,48,160]): the tcgen05 3xTF32 forward is linear in
    its input to fp32 accuracy, conv(2*x1 + x2) == 2*conv(x1) + conv(x2); the TF32 weight gradient is bilinear to TF32
    accuracy; the 16->1 head (exact fp32) likewise."""
    gen = torch.Generator(device=DEV).manual_seed(8)
    shape = (6, 16, 96, 48, 160)
    cl = torch.channels_last_3d
    x1 = torch.randn(shape, device=DEV, generator=gen).contiguous(memory_format=cl)
    x2 = torch.randn(shape, device=DEV, generator=gen).contiguous(memory_format=cl)
    w = 0.1 * torch.randn(16, 16, 3, 3, 3, device=DEV, generator=gen)
    lhs = ops.c16c16_conv_tc(2.0 * x1 + x2, w, 0, 3)
    rhs = 2.0 * ops.c16c16_conv_tc(x1, w, 0, 3) + ops.c16c16_conv_tc(x2, w, 0, 3)
    scale = float(rhs.abs().max())
    assert float((lhs - rhs).abs().max()) < 1e-4 * scale
    gw_l = ops.c16c16_wgrad_tc(x2, 2.0 * x1 + x2)
    gw_r = 2.0 * ops.c16c16_wgrad_tc(x2, x1) + ops.c16c16_wgrad_tc(x2, x2)
    assert float((gw_l - gw_r).abs().max()) < 5e-3 * float(gw_r.abs().max())
    w1 = 0.1 * torch.randn(1, 16, 3, 3, 3, device=DEV, generator=gen)
    yl = ops.conv3d_c16_to_1(2.0 * x1 + x2, w1)
    yr = 2.0 * ops.conv3d_c16_to_1(x1, w1) + ops.conv3d_c16_to_1(x2, w1)
    assert yl.shape == (6, 1, 96, 48, 160)
    assert float((yl - yr).abs().max()) < 1e-4 * float(yr.abs().max())


# ---------------------------------------------------------------------------------------------- K5 + K6 (op level)
def _photo_case(H, W, smooth):
    c = C.case_warp(B=2, H=H, W=W)
    if smooth:                       # image-like statistics at the benchmark resolution
        c["img"] = C.smooth_noise((2, 3, H, W), 73, factor=8, normal=False).clamp(0, 1)
        c["depth"] = 1.0 + 5 * C.smooth_noise((2, 1, H, W), 74, factor=16, normal=False).clamp(0, 1)
    c["tgt"] = c["img"].flip(0).contiguous()
    return c


@pytest.mark.parametrize("ssim_w", [0.85, 0.0], ids=["ssim", "l1only"])
@pytest.mark.parametrize("size,smooth", [((16, 24), False), ((192, 640), True)], ids=["16x24", "192x640"])
def test_photometric_matches_oracle_and_golden(ops, gold, size, smooth, ssim_w):
    """mvd_photometric_fwd/bwd (backproject -> project -> border bilinear warp -> 3x3 SSIM + L1) against the oracle's
    warp_image + reprojection_loss (movedepth/trainer.py:519-550, layers.py:646-677) on identical inputs: loss map, warped
    image, d loss / d depth and d loss / d T.  The second pose of the case throws a band of pixels outside the source
    image (border clamp, zero derivative there).  Bars: 2e-4 abs on colours / losses in [0,1] (sampling coordinates differ
    by <= 6e-5 px from the reference's normalise / unnormalise round trip); gradients within 2e-3 of their scale on
    >= 99.5 % of the pixels (a pixel whose sample sits within 1e-4 px of a cell edge may take the neighbouring cell)."""
    H, W = size
    c = _photo_case(H, W, smooth)
    d_o, T_o = c["depth"].clone().requires_grad_(True), c["T"].clone().requires_grad_(True)
    img_o, _ = OL.warp_image(c["img"], d_o, c["K"], c["invK"], T_o)
    loss_o = OL.reprojection_loss(img_o, c["tgt"], ssim_w, no_ssim=(ssim_w == 0))
    gl = torch.rand(loss_o.shape, generator=torch.Generator().manual_seed(17))
    (loss_o * gl).sum().backward()
    d, T = g(c["depth"]).requires_grad_(True), g(c["T"]).requires_grad_(True)
    loss, warped = ops.photometric_loss(d, g(c["img"]), g(c["tgt"]), g(c["K"]), g(c["invK"]), T, ssim_w)
    (loss * g(gl)).sum().backward()
    torch.testing.assert_close(warped.cpu(), img_o.detach(), atol=2e-4, rtol=0)
    torch.testing.assert_close(loss.detach().cpu(), loss_o.detach(), atol=2e-4, rtol=0)
    if size == (16, 24):             # the reference's own outputs for this case
        np.testing.assert_allclose(warped.cpu().numpy(), gold["warp_img"], atol=2e-4, rtol=0)
        if ssim_w:
            np.testing.assert_allclose(loss.detach().cpu().numpy(), gold["reproj_loss"], atol=2e-4, rtol=0)
    gd, gd_o = d.grad.cpu(), d_o.grad
    err = (gd - gd_o).abs() / float(gd_o.abs().max())
    assert float((err < 2e-3).float().mean()) > 0.995, float((err < 2e-3).float().mean())
    assert float(err.max()) < 0.5                                        # and no wild outlier
    assert 0.0 < float((gd_o == 0).float().mean()) < 0.9                 # some samples fall outside the image: clamped, zero gradient
    torch.testing.assert_close(T.grad.cpu()[:, :3], T_o.grad[:, :3], atol=5e-3 * float(T_o.grad.abs().max()), rtol=5e-3)


def test_photometric_identity_is_ssim_l1_of_the_unwarped_image(ops, gold):
    c = C.case_images()
    got = ops.photometric_identity(g(c["x"]), g(c["y"]), 0.85)
    want = OL.reprojection_loss(c["x"], c["y"], 0.85)
    torch.testing.assert_close(got.cpu(), want, atol=1e-6, rtol=1e-5)
    want_g = 0.85 * torch.from_numpy(gold["ssim"]).mean(1, True) + 0.15 * (c["y"] - c["x"]).abs().mean(1, True)
    torch.testing.assert_close(got.cpu(), want_g, atol=1e-6, rtol=1e-5)
    assert float(ops.photometric_identity(g(c["x"]), g(c["x"]), 0.85).abs().max()) == 0.0        # SSIM(x, x) == 0 exactly


# ---------------------------------------------------------------------------------------------- loss glue
@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_disp_to_depth_full_matches_golden_and_autograd(ops, gold, s):
    """bilinear upsampling (align_corners=False) + disp_to_depth in one kernel vs the reference's own composition
    (tests/golden/ops.npz: up_depth_s*) and the oracle's autograd."""
    c = C.case_disp_pyramid()
    disp = g(c["disp"][s]).requires_grad_(True)
    depth = ops.disp_to_depth_full(disp, c["H"], c["W"], 0.1, 100.0)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), gold["up_depth_s%d" % s], rtol=2e-6, atol=0)
    (depth * g(c["gdepth"])).sum().backward()
    do = c["disp"][s].clone().requires_grad_(True)
    (OL.upsampled_depth(do, c["H"], c["W"], 0.1, 100.0) * c["gdepth"]).sum().backward()
    torch.testing.assert_close(disp.grad.cpu(), do.grad, atol=1e-6 * float(do.grad.abs().max()), rtol=1e-5)


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_smooth_loss_matches_golden_and_autograd(ops, gold, s):
    """mean normalisation + edge-aware smoothness (trainer.py:712-714, layers.py:630-643) vs the reference's value and the
    oracle's autograd; normalize=False is the public get_smooth_loss."""
    import torch.nn.functional as F
    from movedepth_b200 import layers as PL
    c = C.case_disp_pyramid()
    img = F.interpolate(c["img"], [c["H"] // 2 ** s, c["W"] // 2 ** s], mode="area") if s else c["img"]
    for normalize in (True, False):
        disp = g(c["disp"][s]).requires_grad_(True)
        loss = ops.smooth_loss(disp, g(img), normalize=normalize) if normalize else PL.get_smooth_loss(disp, g(img))
        do = c["disp"][s].clone().requires_grad_(True)
        want = OL.normalized_smooth_loss(do, img) if normalize else OL.smooth_loss(do, img)
        if normalize:
            np.testing.assert_allclose(loss.detach().cpu().numpy(), gold["smooth_norm_s%d" % s], rtol=1e-5)
        torch.testing.assert_close(loss.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-7)
        (loss * 3.0).backward()
        (want * 3.0).backward()
        torch.testing.assert_close(disp.grad.cpu(), do.grad, atol=2e-6 * float(do.grad.abs().max()), rtol=1e-4)
    c = C.case_images()
    np.testing.assert_allclose(PL.get_smooth_loss(g(c["disp"]), g(c["x"])).cpu().numpy(), gold["smooth"], rtol=1e-5)


def test_masked_smooth_l1_matches_golden_and_autograd(ops, gold):
    """masked-augmentation consistency (trainer.py:398-400): selection by the resized box mask, mean smooth-L1, weight 100."""
    c = C.case_masked()
    for i, xy in enumerate(c["boxes"]):
        a, b = g(c["a"]).requires_grad_(True), g(c["b"]).requires_grad_(True)
        box = torch.tensor(xy, dtype=torch.int64, device=DEV)
        loss = ops.masked_smooth_l1(a, b, box, c["H"], c["W"], c["H"] // 3, c["W"] // 3, 100.0)
        np.testing.assert_allclose(loss.detach().cpu().numpy(), gold["masked_loss_%d" % i], rtol=1e-5)
        ao, bo = c["a"].clone().requires_grad_(True), c["b"].clone().requires_grad_(True)
        _, m = OL.box_mask(torch.ones(c["B"], 3, c["H"], c["W"]), (c["H"] // 3, c["W"] // 3), xy)
        want = OL.masked_consistency(ao, bo, m, 100.0)
        (loss * 0.5).backward()
        (want * 0.5).backward()
        torch.testing.assert_close(a.grad.cpu(), ao.grad, atol=1e-7, rtol=1e-5)
        torch.testing.assert_close(b.grad.cpu(), bo.grad, atol=1e-7, rtol=1e-5)
        sel = torch.from_numpy(gold["masked_sel_%d" % i])
        assert torch.equal(a.grad.cpu() != 0, sel & ((c["a"] - c["b"]) != 0))


# ---------------------------------------------------------------------------------------------- peer-memory all-reduce
def test_peer_allreduce_two_ranks_on_one_gpu(ops):
    """csrc/peer.cu (the SyncBatchNorm statistics exchange): two 'ranks' = two streams of this process, each with its own
    symmetric buffer on the same GPU, publish into each other's buffers, raise the epoch flags and reduce in rank order.
    Both ranks must hold the bitwise identical sum, over many epochs (double-buffered by epoch parity) and vector lengths."""
    import ctypes
    L = ops._lib.lib()
    world, nmax = 2, 256
    nbytes = L.mvd_peer_allreduce_buffer_bytes(world, nmax)
    bufs = [torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=DEV) for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device=DEV).manual_seed(33)
    torch.cuda.synchronize()
    for it in range(40):
        n = (16, 64, 256, 1)[it % 4]
        vs = [torch.randn(n, dtype=torch.float64, device=DEV, generator=gen) for _ in range(world)]
        outs = [torch.empty(n, dtype=torch.float64, device=DEV) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            rc = L.mvd_peer_allreduce_f64(ctypes.c_void_p(vs[r].data_ptr()), ctypes.c_void_p(outs[r].data_ptr()), n,
                                          ctypes.c_void_p(ptrs.data_ptr()), r, world, nmax, ctypes.c_void_p(streams[r].cuda_stream))
            assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1])
        assert torch.equal(outs[0], vs[0] + vs[1])
    for b in bufs:                                   # [1] of the header = epoch of a timed-out wait (0: none)
        assert int(b.view(torch.int64)[1]) == 0 and int(b.view(torch.int64)[0]) == 40


# ---------------------------------------------------------------------------------------------- DepthDecoder glue
@pytest.mark.parametrize("shape,C2,up,act,bias", [((2, 16, 6, 10), 0, 1, True, True), ((2, 32, 5, 7), 64, 2, True, True),
                                                   ((1, 512, 2, 4), 0, 1, False, False), ((2, 2048, 2, 3), 0, 1, False, False), ((1, 8, 3, 3), 0, 1, True, True),
                                                   ((1, 256, 3, 4), 256, 2, True, True)],
                         ids=["elu-pad", "elu-up-cat-pad", "plain-pad", "plain-pad-r50", "3x3", "c256"])
def test_decoder_prep_matches_the_tensor_composition(ops, shape, C2, up, act, bias):
    """mvd_decoder_prep_{fwd,bwd}: ReflectionPad2d(1)(cat(nearest_up(ELU(z + bias)), skip)) and its 3xTF32 split in one kernel
    vs the torch ops the reference's DepthDecoder strings together (depth_decoder.py:72-101), forward and all gradients."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(23)
    B, C1, h, w = shape
    z = torch.randn(shape, generator=gen)
    b = 0.3 * torch.randn(C1, generator=gen) if bias else None
    skip = torch.randn(B, C2, h * up, w * up, generator=gen) if C2 else None
    gout = torch.randn(B, C1 + C2, h * up + 2, w * up + 2, generator=gen)

    def compose(z, b, skip):
        t = z + b.view(1, -1, 1, 1) if b is not None else z
        t = F.elu(t) if act else t
        if up == 2:
            t = F.interpolate(t, scale_factor=2, mode="nearest")
        if skip is not None:
            t = torch.cat([t, skip], 1)
        return F.pad(t, (1, 1, 1, 1), mode="reflect")

    leaves = [t.double().requires_grad_(True) if t is not None else None for t in (z, b, skip)]
    want = compose(*leaves)
    (want * gout.double()).sum().backward()
    cl = torch.channels_last
    zg = g(z).contiguous(memory_format=cl).requires_grad_(True)
    bg = g(b).requires_grad_(True) if bias else None
    sg = g(skip).contiguous(memory_format=cl).requires_grad_(True) if C2 else None
    xp, x3 = ops.decoder_prep(zg, bg, sg, act=act, up=up, want_split=True)
    torch.testing.assert_close(xp.cpu(), want.detach().float(), atol=1e-6, rtol=1e-6)
    C = C1 + C2
    hi, lo, hi2 = x3[:, :C], x3[:, C:2 * C], x3[:, 2 * C:]
    assert torch.equal(hi, hi2) and torch.equal(hi + lo, xp)                       # exact split
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0                   # hi has a 10-bit mantissa
    (xp * g(gout)).sum().backward()
    torch.testing.assert_close(zg.grad.cpu(), leaves[0].grad.float(), atol=1e-5, rtol=1e-5)
    if bias:
        torch.testing.assert_close(bg.grad.cpu(), leaves[1].grad.float(), atol=1e-4, rtol=1e-4)
    if C2:
        torch.testing.assert_close(sg.grad.cpu(), leaves[2].grad.float(), atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("arch", [18, 50])
def test_fused_depth_decoder_matches_the_oracle_decoder(ops, arch):
    """DepthDecoder on the glue kernel + pre-padded convolutions (fp32 policy) vs the oracle's module-by-module decoder on the
    CPU: the four disparities and the gradients of every parameter and of the encoder features."""
    from movedepth_b200 import networks as PN, precision as PR
    from oracle import networks as ON
    PR.set_policy("fp32")
    gen = torch.Generator().manual_seed(31)
    ch = [64, 64, 128, 256, 512] if arch == 18 else [64, 256, 512, 1024, 2048]
    feats = [torch.randn(2, c, 32 >> i, 64 >> i, generator=gen) for i, c in enumerate(ch)]
    a, b = PN.DepthDecoder(np.array(ch)), ON.DepthDecoder(np.array(ch))
    fill_deterministic(a)
    fill_deterministic(b)
    a.to(DEV)
    fa = [g(f).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
    fb = [f.clone().requires_grad_(True) for f in feats]
    oa, ob = a(fa), b(fb)
    loss_a = sum((oa[("disp", s)] * (s + 1)).sum() for s in range(4))
    loss_b = sum((ob[("disp", s)] * (s + 1)).sum() for s in range(4))
    loss_a.backward()
    loss_b.backward()
    for s in range(4):
        torch.testing.assert_close(oa[("disp", s)].detach().cpu(), ob[("disp", s)].detach(), atol=2e-5, rtol=1e-4)
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        torch.testing.assert_close(pa.grad.cpu(), pb.grad, atol=2e-4 * float(pb.grad.abs().max()) + 1e-7, rtol=1e-3, msg=n)
    for x, y in zip(fa, fb):
        torch.testing.assert_close(x.grad.cpu(), y.grad, atol=2e-4 * float(y.grad.abs().max()), rtol=1e-3)


def test_split_tf32_vectorised_and_scalar_paths(ops):
    """mvd_split_tf32: x = hi + lo exactly, hi rounded to TF32; C % 4 == 0 takes the 16-byte path, C = 3 the scalar one."""
    from movedepth_b200 import precision as PR
    for shape in ((2, 16, 5, 7), (1, 3, 6, 9), (2, 8, 3, 4, 5)):
        fmt = torch.channels_last_3d if len(shape) == 5 else torch.channels_last
        x = torch.randn(shape, device=DEV).contiguous(memory_format=fmt)
        for weight in (False, True):
            y = PR._split_dim1(x, weight)
            C = shape[1]
            hi = PR.tf32_round(x)
            want = torch.cat([hi, hi, x - hi] if weight else [hi, x - hi, hi], 1)
            assert torch.equal(y, want) and y.shape[1] == 3 * C


def test_batchnorm_reductions_exchange_inside_the_kernel_two_ranks_on_one_gpu(ops):
    """Data-parallel BatchNorm: the last block of mvd_bn_stats / mvd_bn_bwd_reduce all-reduces the 2C fp64 sums over peer
    memory itself (no separate exchange launch).  Two 'ranks' = two streams with their own symmetric buffers on this GPU:
    both must end with the sums over BOTH halves of the batch, and the backward keeps each rank's own sums aside."""
    import ctypes
    L = ops._lib.lib()
    world, nmax, C, M = 2, 2048, 16, 5000
    nbytes = L.mvd_peer_allreduce_buffer_bytes(world, nmax)
    bufs = [torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=DEV) for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device=DEV).manual_seed(35)
    P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
    for it in range(6):
        xs = [torch.randn(M + 100 * r, C, device=DEV, generator=gen) * 2 + 0.5 for r in range(world)]
        gs = [torch.randn(M + 100 * r, C, device=DEV, generator=gen) for r in range(world)]
        sums = [torch.empty(2 * C + 1, dtype=torch.float64, device=DEV) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            rc = L.mvd_bn_stats(P(xs[r]), xs[r].shape[0], C, P(sums[r]), P(ptrs), r, world, nmax, ctypes.c_void_p(streams[r].cuda_stream))
            assert rc == 0
        torch.cuda.synchronize()
        allx = torch.cat(xs).double()
        want = torch.cat([allx.sum(0), (allx * allx).sum(0)])
        assert torch.equal(sums[0][:2 * C], sums[1][:2 * C])
        torch.testing.assert_close(sums[0][:2 * C], want, rtol=1e-6, atol=1e-6)
        # backward reduction: sum g and sum g * xhat (stats: mean 0.5, invstd 0.5, scale / shift unused without ReLU)
        stats = torch.cat([torch.full((C,), 0.5), torch.full((C,), 0.5), torch.ones(C), torch.zeros(C)]).to(DEV)
        sums2 = [torch.empty(2 * C + 1, dtype=torch.float64, device=DEV) for _ in range(world)]
        local = [torch.empty(2 * C, dtype=torch.float64, device=DEV) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            rc = L.mvd_bn_bwd_reduce(P(gs[r]), P(xs[r]), P(None), P(stats), P(sums2[r]), P(local[r]), xs[r].shape[0], C, 0, P(ptrs), r,
                                     world, nmax, ctypes.c_void_p(streams[r].cuda_stream))
            assert rc == 0
        torch.cuda.synchronize()
        mine = [torch.cat([gs[r].double().sum(0), (gs[r].double() * (xs[r].double() - 0.5) * 0.5).sum(0)]) for r in range(world)]
        for r in range(world):
            torch.testing.assert_close(local[r], mine[r], rtol=1e-5, atol=1e-5)
        assert torch.equal(sums2[0][:2 * C], sums2[1][:2 * C])
        torch.testing.assert_close(sums2[0][:2 * C], mine[0] + mine[1], rtol=1e-5, atol=1e-5)
    for b in bufs:
        assert int(b.view(torch.int64)[1]) == 0 and int(b.view(torch.int64)[0]) == 12


def test_one_kernel_batchnorm_exchanges_at_its_grid_barrier_two_ranks_on_one_gpu(ops):
    """SyncBatchNorm in one launch per direction: the last block to reach the grid barrier of mvd_bn_fwd_fused /
    mvd_bn_bwd_fused all-reduces the sums over peer memory before it opens the barrier.  Two 'ranks' = two streams on this GPU
    (their two barrier kernels are resident together and wait for each other): both must normalise with the statistics of
    BOTH halves of the batch, the backward keeps each rank's own sums aside, and the workspaces come back zeroed."""
    import ctypes
    L = ops._lib.lib()
    world, nmax, C = 2, 2048, 32
    nbytes = L.mvd_peer_allreduce_buffer_bytes(world, nmax)
    bufs = [torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=DEV) for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    wss = [torch.zeros(L.mvd_bn_workspace_doubles(), dtype=torch.float64, device=DEV) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device=DEV).manual_seed(36)
    P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
    S = lambda r: ctypes.c_void_p(streams[r].cuda_stream)
    w = 1 + 0.2 * torch.randn(C, device=DEV, generator=gen)
    b = 0.1 * torch.randn(C, device=DEV, generator=gen)
    for it, M in enumerate((300, 70000, 5000)):                                   # one block / the full grid / in between
        Ms = [M, M + 123]
        xs = [torch.randn(Ms[r], C, device=DEV, generator=gen) * 2 + 0.5 for r in range(world)]
        gs = [torch.randn(Ms[r], C, device=DEV, generator=gen) for r in range(world)]
        count = float(sum(Ms))
        ys = [torch.empty_like(x) for x in xs]
        stats = [torch.empty(4 * C, device=DEV) for _ in range(world)]
        rm = [torch.zeros(C, device=DEV) for _ in range(world)]
        rv = [torch.ones(C, device=DEV) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            assert L.mvd_bn_fwd_fused(P(xs[r]), P(None), P(w), P(b), P(rm[r]), P(rv[r]), P(None), 0.1, 1e-5, count, P(stats[r]),
                                      P(ys[r]), Ms[r], C, 1, P(wss[r]), P(ptrs), r, world, nmax, S(r)) == 0
        torch.cuda.synchronize()
        allx = torch.cat(xs).double()
        mean, var = allx.mean(0), allx.var(0, unbiased=False)
        assert torch.equal(stats[0], stats[1])
        torch.testing.assert_close(stats[0][:C], mean.float(), atol=1e-6, rtol=1e-6)
        torch.testing.assert_close(rv[0], (0.9 + 0.1 * allx.var(0, unbiased=True)).float(), atol=1e-6, rtol=1e-6)
        for r in range(world):
            want = torch.relu((xs[r].double() - mean) / torch.sqrt(var + 1e-5) * w.double() + b.double())
            torch.testing.assert_close(ys[r], want.float(), atol=2e-5, rtol=2e-5)
        gxs = [torch.empty_like(x) for x in xs]
        local = [torch.empty(2 * C, dtype=torch.float64, device=DEV) for _ in range(world)]
        for r in range(world):
            assert L.mvd_bn_bwd_fused(P(gs[r]), P(xs[r]), P(None), P(stats[r]), P(w), count, P(gxs[r]), P(None), P(None), P(None),
                                      P(local[r]), Ms[r], C, 1, P(wss[r]), P(ptrs), r, world, nmax, S(r)) == 0
        torch.cuda.synchronize()
        xo = torch.cat(xs).double().requires_grad_(True)
        wo = w.double().requires_grad_(True)
        yo = torch.relu((xo - xo.mean(0)) / torch.sqrt(xo.var(0, unbiased=False) + 1e-5) * wo + b.double())
        (yo * torch.cat(gs).double()).sum().backward()
        got = torch.cat(gxs)
        torch.testing.assert_close(got, xo.grad.float(), atol=1e-4 * float(xo.grad.abs().max()), rtol=1e-4)
        torch.testing.assert_close((local[0][C:] + local[1][C:]).float(), wo.grad.float(), atol=1e-3, rtol=1e-4)
        for r in range(world):
            assert int(wss[r].view(torch.int64).abs().sum()) == 0


def test_pose_matrix_kernel_matches_the_reference_golden(ops, gold):
    """mvd_pose_matrix_fwd vs the reference's own transformation_from_parameters outputs (tests/golden/ops.npz: T_fwd, T_inv,
    movedepth/layers.py:412-429), incl. the zero rotation that exercises the angle + 1e-7 guard."""
    c = C.case_pose()
    for key, invert in (("T_fwd", False), ("T_inv", True)):
        got = ops.pose_matrix(g(c["aa"]), g(c["tr"]), invert).cpu()
        torch.testing.assert_close(got, torch.from_numpy(gold[key]), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("invert", [False, True], ids=["forward", "inverted"])
def test_pose_matrix_kernel_matches_the_oracle(ops, invert):
    """mvd_pose_matrix_fwd/bwd vs oracle.layers.transformation_from_parameters (movedepth/layers.py:412-429) in fp64: the 4x4
    transform, and the gradients of a random linear functional w.r.t. axis-angle and translation (incl. a zero rotation,
    where torch's norm backward defines the sub-gradient as 0)."""
    gen = torch.Generator().manual_seed(51)
    aa = torch.randn(7, 1, 3, generator=gen) * 0.3
    tr = torch.randn(7, 1, 3, generator=gen)
    aa[3] = 0.0
    aa[5] *= 1e-3
    gM = torch.randn(7, 4, 4, generator=gen)
    ao, to = aa.double().requires_grad_(True), tr.double().requires_grad_(True)
    Mo = OL.transformation_from_parameters(ao, to, invert)
    (Mo * gM.double()).sum().backward()
    ag, tg = g(aa).requires_grad_(True), g(tr).requires_grad_(True)
    M = ops.pose_matrix(ag, tg, invert)
    (M * g(gM)).sum().backward()
    torch.testing.assert_close(M.detach().cpu(), Mo.detach().float(), atol=2e-6, rtol=1e-5)
    keep = [i for i in range(7) if i != 5]
    torch.testing.assert_close(ag.grad.cpu()[keep], ao.grad.float()[keep], atol=1e-5, rtol=1e-4)
    # item 5 is a 3e-4 rad rotation: 1 - cos(theta) keeps ~3 significant digits in fp32 (in the reference's arithmetic too)
    torch.testing.assert_close(ag.grad.cpu()[5], ao.grad.float()[5], atol=5e-4, rtol=5e-3)
    torch.testing.assert_close(tg.grad.cpu(), to.grad.float(), atol=1e-5, rtol=1e-4)
    from movedepth_b200 import layers as PL
    assert torch.equal(PL.transformation_from_parameters(g(aa), g(tr), invert), M.detach())     # the public function takes the kernel


# ---------------------------------------------------------------------------------------------- skinny 2-D convolutions
@pytest.mark.parametrize("cfg", [(3, 8, 3, 1), (8, 8, 3, 1), (8, 16, 5, 2), (16, 16, 3, 1), (16, 16, 3, 1, 0), (16, 4, 3, 1), (16, 4, 3, 1, 0)],
                         ids=["3-8", "8-8", "8-16-k5s2", "16-16", "16-16-valid", "16-4", "16-4-valid"])
@pytest.mark.parametrize("hw", [(16, 32), (13, 45), (48, 160)], ids=["aligned", "ragged", "multi-tile"])
def test_conv2d_small_matches_torch_conv2d(ops, cfg, hw):
    """csrc/conv2d_small.cu (FPN4's conv0 / conv1 stages, UncertNet's 8->8 layer, the DepthDecoder's finest stage on pre-padded
    inputs = "valid"): exact-fp32 direct forward, data gradient (flipped-filter forward kernel for stride 1, parity gather for
    stride 2) and weight gradient vs torch's CPU conv2d in fp64; 1e-5 of the output scale (fp32 accumulation of <= 400 products)."""
    import torch.nn.functional as F
    cin, cout, k, s = cfg[:4]
    pad = cfg[4] if len(cfg) > 4 else k // 2
    H, W = hw
    if s == 2:
        H, W = H & ~1, W & ~1
    gen = torch.Generator().manual_seed(41)
    x = torch.randn(2, cin, H, W, generator=gen)
    w = torch.randn(cout, cin, k, k, generator=gen) * 0.2
    xo, wo = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yo = F.conv2d(xo, wo, stride=s, padding=pad)
    gy = torch.randn(yo.shape, generator=gen)
    (yo * gy.double()).sum().backward()
    need_gx = cin > 3
    xg = g(x).contiguous(memory_format=torch.channels_last).requires_grad_(need_gx)
    wg = g(w).requires_grad_(True)
    assert ops.conv2d_small_supported(cin, cout, k, s)
    y = ops.conv2d_small(xg, wg, k, s, pad=pad)
    assert y.shape == yo.shape
    (y * g(gy)).sum().backward()
    pairs = [(y, yo), (wg.grad, wo.grad)] + ([(xg.grad, xo.grad)] if need_gx else [])
    for got, want in pairs:
        want = want.detach().float()
        torch.testing.assert_close(got.detach().cpu(), want, atol=1e-5 * float(want.abs().max()), rtol=1e-5)


def test_fpn4_routes_its_skinny_layers_through_the_direct_kernels(ops):
    """FPN4 (fp32 policy) with the direct kernels vs the oracle FPN4 on the CPU: both outputs and every parameter gradient;
    the launch counter proves the conv0 / conv1 stages took the direct path."""
    from movedepth_b200 import networks as PN, precision as PR
    from oracle import networks as ON
    PR.set_policy("fp32")
    gen = torch.Generator().manual_seed(43)
    img = torch.rand(2, 3, 64, 96, generator=gen)
    a, b = PN.FPN4(8, 2), ON.FPN4(8, 2)
    fill_deterministic(a)
    fill_deterministic(b)
    a.to(DEV)
    n0 = ops.launch_counter["n"]
    fa, ca = a(g(img).contiguous(memory_format=torch.channels_last))
    assert ops.launch_counter["n"] - n0 >= 5                      # five skinny convolutions + the BatchNorm kernels
    fb, cb = b(img)
    torch.testing.assert_close(fa.detach().cpu(), fb.detach(), atol=1e-4 * float(fb.abs().max()), rtol=1e-4)
    torch.testing.assert_close(ca.detach().cpu(), cb.detach(), atol=1e-4 * float(cb.abs().max()), rtol=1e-4)
    gf, gc = torch.randn(fb.shape, generator=gen), torch.randn(cb.shape, generator=gen)
    ((fa * g(gf)).sum() + (ca * g(gc)).sum()).backward()
    ((fb * gf).sum() + (cb * gc).sum()).backward()
    # The seeds gf / gc are white noise, so every per-channel gradient is a heavily cancelling sum over 12 k pixels: ONE ReLU
    # mask flipping under 1e-7 summation-order noise moves a BatchNorm bias gradient by ~1 % of its value (measured 1.3 %).
    # The kernels themselves are pinned to 1e-5 by the op-level test above; this chained test bounds gross errors only.
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        err = float((pa.grad.cpu() - pb.grad).abs().max() / (pb.grad.abs().max() + 1e-12))
        assert err < 5e-2, (n, err)


def test_conv0_epilogue_accumulates_the_batchnorm_statistics(ops):
    """The tcgen05 16->16 conv can add the per-channel sum / sum of squares of its output in the epilogue (fused BatchNorm
    statistics): they must equal the sums of the tensor it wrote, and ConvBnReLU3D must produce the same activations and
    running statistics with and without the fusion."""
    from movedepth_b200 import networks as PN, precision as PR, norm as NM
    from movedepth_b200.networks.resnet_encoder import ConvBnReLU3D
    gen = torch.Generator(device=DEV).manual_seed(47)
    x = torch.randn(2, 16, 13, 20, 45, device=DEV, generator=gen).contiguous(memory_format=torch.channels_last_3d)
    w = 0.1 * torch.randn(16, 16, 3, 3, 3, device=DEV, generator=gen)
    y, sums = ops.conv3d_c16_to_16_with_stats(x, w, 3)
    yd = y.double()
    torch.testing.assert_close(sums[:16], yd.sum((0, 2, 3, 4)), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(sums[16:32], (yd * yd).sum((0, 2, 3, 4)), rtol=1e-5, atol=1e-4)
    # (not bitwise: the taps are issued by four warps into one accumulator, the hardware's accumulation order may differ)
    torch.testing.assert_close(y, ops.conv3d_c16_to_16(x, w, 3), atol=1e-5, rtol=1e-5)
    blocks = []
    for policy in ("3xtf32", "fp32"):                    # fused statistics under the default policy, separate kernel otherwise
        PR.set_policy(policy)
        torch.manual_seed(0)
        blk = ConvBnReLU3D(16, 16).to(DEV)
        with torch.no_grad():
            blk.conv.weight.copy_(w)
        blocks.append((blk, blk(x)))
    PR.set_policy("fp32")
    (b0, y0), (b1, y1) = blocks
    torch.testing.assert_close(y0, y1, atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(b0.bn.running_var, b1.bn.running_var, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(b0.bn.running_mean, b1.bn.running_mean, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("shape", [(2, 64, 24, 40), (1, 8, 7, 9), (2, 16, 1, 5)], ids=["even", "odd", "one-row"])
def test_maxpool3x3s2_matches_torch(ops, shape):
    """ResNet stem MaxPool2d(3, 2, 1): forward values and the gradient routed to the (first) maximum of every window."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(53)
    x = torch.randn(shape, generator=gen)
    x[0, :, 0, :3] = 1.5                                   # ties: the first maximum in scan order takes the gradient
    xo = x.clone().requires_grad_(True)
    yo = F.max_pool2d(xo, 3, 2, 1)
    gy = torch.randn(yo.shape, generator=gen)
    (yo * gy).sum().backward()
    xg = g(x).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = ops.maxpool3x3s2(xg)
    (y * g(gy)).sum().backward()
    assert torch.equal(y.detach().cpu(), yo.detach())
    torch.testing.assert_close(xg.grad.cpu(), xo.grad, atol=1e-6, rtol=1e-6)


# ---------------------------------------------------------------------------------------------- device data pipeline
def test_device_pipeline_is_bit_exact_to_the_reference_preprocess(ops):
    """movedepth_b200.datapipe.DevicePipeline (Lanczos pyramid, flip, ToTensor, ColorJitter on the GPU) vs the reference's own
    MonoDataset.preprocess output (tests/golden/datapipe.npz): every pixel of every scale identical, with and without the
    flip, and with the jitter parameters drawn from the same torch seed in the same order."""
    from movedepth_b200.datapipe import DevicePipeline
    gold = dict(np.load(os.path.join(GOLD, "datapipe.npz")))
    c = C.case_frames()
    pipe = DevicePipeline(c["H"], c["W"], device=DEV)
    K = [[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    frames = {f: torch.from_numpy(c["frames"][f].copy())[None].to(DEV) for f in (0, -1)}
    plain = pipe(frames, K)
    flipped = pipe(frames, K, flip=[1])
    torch.manual_seed(c["seed"])
    jit = pipe(frames, K, color_aug=[1])
    for f in (0, -1):
        for s in range(4):
            assert np.array_equal(plain[("color", f, s)][0].cpu().numpy(), gold["plain_color_%d_%d" % (f, s)])
            assert np.array_equal(plain[("color_aug", f, s)][0].cpu().numpy(), gold["plain_color_%d_%d" % (f, s)])
            assert np.array_equal(flipped[("color", f, s)][0].cpu().numpy(), gold["flip_color_%d_%d" % (f, s)])
            assert np.array_equal(jit[("color", f, s)][0].cpu().numpy(), gold["jitter_color_%d_%d" % (f, s)])
            assert np.array_equal(jit[("color_aug", f, s)][0].cpu().numpy(), gold["jitter_color_aug_%d_%d" % (f, s)]), (f, s)
    assert plain[("K", 2)].shape == (1, 4, 4) and abs(float(plain[("K", 2)][0, 0, 0]) - 0.58 * 24) < 1e-5


def test_device_pipeline_matches_the_oracle_at_kitti_size(ops):
    """375x1242 'decoded' frames -> 192x640 pyramid, batch of 3 with mixed flips and per-image jitter (random orders, all four
    operations, extrapolating and interpolating factors): bit-exact against oracle/datapipe.py (pinned to Pillow)."""
    from movedepth_b200.datapipe import DevicePipeline
    from oracle import datapipe as OD
    rng = np.random.default_rng(5)
    B, Hn, Wn, H, W = 3, 375, 1242, 192, 640
    walk = np.cumsum(rng.normal(0, 5, (B, Hn, Wn, 3)), 2) + rng.normal(0, 8, (B, Hn, Wn, 3))
    native = np.clip(walk + 120, 0, 255).astype(np.uint8)
    pipe = DevicePipeline(H, W, device=DEV)
    dev = torch.from_numpy(native).to(DEV)
    flip = [0, 1, 0]
    img = dev
    for s in range(4):
        img = pipe.resize(img, H >> s, W >> s, torch.tensor(flip, dtype=torch.uint8, device=DEV) if s == 0 else None)
        for n in range(B):
            src = native[n][:, ::-1] if flip[n] else native[n]
            want = OD.pyramid(src, H, W)[s] if s < 2 else None
            if want is not None:
                assert np.array_equal(img[n].cpu().numpy(), want), (s, n)
    base = pipe.resize(dev, H, W)
    orders = [[0, 1, 2, 3], [3, 2, 1, 0], [2, 0, 3, 1]]
    factors = [[0.8, 1.2, 0.93, -0.1], [1.17, 0.85, 1.2, 0.07], [1.0, 1.0, 0.8, 0.0]]
    got = pipe.color_jitter(base, orders, factors, [1, 1, 1]).cpu().numpy()
    for n in range(B):
        assert np.array_equal(got[n], OD.color_jitter(base[n].cpu().numpy(), orders[n], factors[n])), n
    t = pipe.to_tensor(base)
    assert np.array_equal(t[1].cpu().numpy(), OD.to_tensor(base[1].cpu().numpy()))
