"""Host-side logic of the product package, on CPU: checkpoint key layout, parameter counts, the
plain-tensor glue in layers.py against the oracle, option parsing, flat arenas, and the
world_size-2 gradient exchange (gloo)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _cases as C
from oracle import layers as OL
from oracle import networks as ON
from movedepth_b200 import layers as PL
from movedepth_b200 import networks as PN
from movedepth_b200.options import MonodepthOptions
from movedepth_b200.trainer import FlatArena, SyntheticKITTI, kitti_intrinsics


def _pairs():
    enc_o, enc_p = ON.ResnetEncoder(18, False), PN.ResnetEncoder(18, False)
    yield "mono_encoder", enc_o, enc_p, 11176512
    yield "mono_depth", ON.DepthDecoder(enc_o.num_ch_enc, range(4)), PN.DepthDecoder(enc_p.num_ch_enc, range(4)), 3152724
    yield "pose_encoder", ON.ResnetEncoder(18, False, 2), PN.ResnetEncoder(18, False, num_input_images=2), 11185920
    yield "pose", ON.PoseDecoder(enc_o.num_ch_enc, 1, 2), PN.PoseDecoder(enc_p.num_ch_enc, 1, 2), 1314572
    yield "mvs_encoder", ON.FPN4(8, 2), PN.FPN4(base_channels=8, scale=2), 186008
    yield "reg3d", ON.Reg3d(16, 16, 3), PN.reg3d(16, 16, 3), 1169712
    yield "mask_cnn", ON.UncertNet(), PN.UncertNet(), 752
    yield "up", ON.ConvexUpsampleLayer(32, 2), PL.convex_upsample_layer(32, 2), 27648


def test_state_dict_keys_and_param_counts_match_reference_layout():
    for name, o, p, count in _pairs():
        so, sp = o.state_dict(), p.state_dict()
        assert list(so.keys()) == list(sp.keys()), name
        assert all(so[k].shape == sp[k].shape for k in so), name
        assert sum(t.numel() for t in p.parameters()) == count, name


def test_r50_encoder_channels():
    e = PN.ResnetEncoder(50, False)
    assert list(e.num_ch_enc) == [64, 256, 512, 1024, 2048]
    assert sum(t.numel() for t in e.parameters()) == 23508032


def test_networks_forward_equal_oracle_on_cpu():
    torch.manual_seed(0)
    x = torch.rand(1, 3, 64, 96)
    for name, o, p, _ in _pairs():
        p.load_state_dict(o.state_dict())
        o.eval(), p.eval()
    pairs = {n: (o, p) for n, o, p, _ in _pairs()}
    o, p = ON.FPN4(8, 2).eval(), PN.FPN4(8, 2).eval()
    p.load_state_dict(o.state_dict())
    for a, b in zip(o(x), p(x)):
        assert torch.equal(a, b)
    o, p = ON.Reg3d(16, 16, 3).eval(), PN.reg3d(16, 16, 3).eval()
    p.load_state_dict(o.state_dict())
    v = torch.rand(1, 8, 16, 16, 24)
    assert torch.equal(o(v), p(v))
    assert torch.equal(o(v), p.forward_volume(v.permute(0, 2, 1, 3, 4)))
    o, p = ON.UncertNet().eval(), PN.UncertNet().eval()
    p.load_state_dict(o.state_dict())
    e = torch.rand(1, 1, 16, 24)
    assert torch.equal(o(e), p(e))
    eo, ep = ON.ResnetEncoder(18, False).eval(), PN.ResnetEncoder(18, False).eval()
    ep.load_state_dict(eo.state_dict())
    do, dp = ON.DepthDecoder(eo.num_ch_enc).eval(), PN.DepthDecoder(ep.num_ch_enc).eval()
    dp.load_state_dict(do.state_dict())
    a, b = do(eo(x)), dp(ep(x))
    for s in range(4):
        assert torch.equal(a[("disp", s)], b[("disp", s)])
    po, pp = ON.PoseDecoder(eo.num_ch_enc, 1, 2).eval(), PN.PoseDecoder(ep.num_ch_enc, 1, 2).eval()
    pp.load_state_dict(po.state_dict())
    for u, v in zip(po([eo(x)]), pp([ep(x)])):
        assert torch.equal(u, v)


def test_pose_and_disparity_glue_match_oracle():
    c = C.case_pose()
    for inv in (False, True):
        a = OL.transformation_from_parameters(c["aa"], c["tr"], inv)
        b = PL.transformation_from_parameters(c["aa"], c["tr"], inv)
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
    d = torch.rand(2, 1, 4, 4)
    for a, b in zip(OL.disp_to_depth(d, 0.1, 100.0), PL.disp_to_depth(d, 0.1, 100.0)):
        assert torch.equal(a, b)


def test_hypothesis_schedules_and_separable_ratios():
    c = C.case_hypotheses()
    for kind in ("inverse", "linear"):
        assert torch.equal(OL.depth_hypotheses(c["prior"], c["D"], c["fac"], kind=kind),
                           PL.schedule_depth_rangev2(c["prior"], c["D"], c["fac"], kind))
    want = OL.depth_hypotheses(c["prior"], c["D"], c["fac"], z_trans=c["z_trans"])
    assert torch.equal(want, PL.schedule_depth_range_zv2(c["prior"], c["D"], c["fac"], c["z_trans"]))
    s = (c["fac"] * c["z_trans"]).reshape(-1)
    sep = c["prior"] * PL.hypothesis_ratios(c["D"], s, "cpu").view(2, c["D"], 1, 1)
    torch.testing.assert_close(sep, want, rtol=2e-6, atol=0)


def test_ssim_smoothness_projection_modules_match_oracle():
    c = C.case_images()
    torch.testing.assert_close(PL.SSIM()(c["x"], c["y"]), OL.ssim(c["x"], c["y"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(PL.get_smooth_loss(c["disp"], c["x"]), OL.smooth_loss(c["disp"], c["x"]), rtol=1e-6, atol=0)
    w = C.case_warp()
    pts = PL.BackprojectDepth(w["B"], w["H"], w["W"])(w["depth"], w["invK"])
    torch.testing.assert_close(pts, OL.backproject(w["depth"], w["invK"], w["H"], w["W"]), rtol=1e-6, atol=1e-6)
    grid = PL.Project3D(w["B"], w["H"], w["W"])(pts, w["K"], w["T"])
    torch.testing.assert_close(grid, OL.project(pts, w["K"], w["T"], w["H"], w["W"]), rtol=1e-5, atol=1e-6)


def test_options_keep_reference_flags_and_defaults():
    o = MonodepthOptions().parse([])
    assert (o.height, o.width, o.num_depth_bins, o.reg3d_c, o.prior_scale, o.norm_radius) == (192, 640, 16, 16, 2, 1)
    assert o.frame_ids == [0, -1, 1] and o.matching_ids == [0, -1] and o.scales == [0, 1, 2, 3]
    assert (o.depth_bin_fac, o.z_scale, o.ztrans_start_epc, o.ssim_lw, o.mask_lw) == (0.3, 30, 8, 0.85, 10)
    assert o.learning_rate == 1e-4 and o.scheduler_step_size == 15 and o.batch_size == 12
    o = MonodepthOptions().parse("--frame_ids 0 -1 --convex_up --ddp --local_rank 3 --learning_rate 2e-4".split())
    assert o.frame_ids == [0, -1] and o.convex_up and o.ddp and o.local_rank == 3 and o.learning_rate == 2e-4


def test_product_refuses_to_run_without_cuda():
    from movedepth_b200.trainer import Trainer
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    o = MonodepthOptions().parse(["--frame_ids", "0", "-1", "--weights_init", "scratch", "--no_cuda"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        Trainer(o)
    from movedepth_b200 import ops
    c = C.case_convex()
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.convex_upsample(c["depth"], c["mask"], 2)


def test_flat_arena_rehomes_parameters_and_grads():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    before = [p.detach().clone() for p in net.parameters()]
    arena = FlatArena(net.parameters(), "cpu")
    assert arena.numel % 4 == 0 and all(o % 4 == 0 for o in arena.offsets)
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p, b) and p.data_ptr() >= arena.data.data_ptr()
    net(torch.rand(4, 5)).sum().backward()
    g = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert float(arena.grad.abs().sum()) == pytest.approx(float(g.abs().sum()), rel=1e-6)
    arena.data.mul_(0)
    assert all(float(p.abs().sum()) == 0 for p in net.parameters())


def test_synthetic_batches_follow_the_item_schema():
    o = MonodepthOptions().parse(["--frame_ids", "0", "-1", "--height", "64", "--width", "96"])
    item = next(iter(SyntheticKITTI(o, 2, 1, pin=False)))
    for f in (0, -1):
        for s in range(4):
            assert item[("color", f, s)].shape == (2, 3, 64 >> s, 96 >> s)
            assert item[("color_aug", f, s)].shape == (2, 3, 64 >> s, 96 >> s)
    K, iK = kitti_intrinsics(2, 16, 24)
    torch.testing.assert_close(item[("K", 2)], K)
    torch.testing.assert_close(torch.matmul(K, iK), torch.eye(4).repeat(2, 1, 1), atol=1e-5, rtol=0)


def _exchange_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Linear(6, 4)
    arena = FlatArena(net.parameters(), "cpu")
    x = torch.full((3, 6), float(rank + 1))
    net(x).sum().backward()
    work = dist.all_reduce(arena.grad, async_op=True)     # the trainer's exchange: SUM, scaled in the Adam kernel
    work.wait()
    torch.save(arena.grad / world, os.path.join(out, "g%d.pt" % rank))
    dist.destroy_process_group()


def test_gradient_exchange_world_size_2_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)
    # mean over ranks of d/dW sum(Wx+b): x summed over the 3 rows -> 3*(1+2)/2 = 4.5 per weight, 3 per bias
    assert torch.allclose(g0[:24], torch.full((24,), 4.5)) and torch.allclose(g0[24:28], torch.full((4,), 3.0))


def test_3xtf32_split_conv_matches_plain_conv_and_its_gradients():
    """movedepth_b200/precision.py: the split evaluates x_hi*w_hi + x_lo*w_hi + x_hi*w_lo; on CPU (exact fp32
    products) that differs from conv(x, w) only by the dropped x_lo*w_lo term (~2^-22 relative)."""
    from movedepth_b200 import precision as PR
    g = torch.Generator().manual_seed(0)
    cases = [
        (PR.Conv3d(4, 6, 3, stride=2, padding=1, bias=False), torch.nn.Conv3d(4, 6, 3, stride=2, padding=1, bias=False), (2, 4, 8, 8, 10)),
        (PR.ConvTranspose3d(6, 4, 3, stride=2, padding=1, output_padding=1, bias=False),
         torch.nn.ConvTranspose3d(6, 4, 3, stride=2, padding=1, output_padding=1, bias=False), (2, 6, 4, 4, 5)),
        (PR.Conv2d(3, 5, 5, stride=2, padding=2, bias=True), torch.nn.Conv2d(3, 5, 5, stride=2, padding=2, bias=True), (2, 3, 12, 14)),
    ]
    for mine, plain, shape in cases * 2:
        split_bwd = mine.weight.grad is not None         # second visit of each case: split the backward too
        mine.weight.grad = plain.weight.grad = None
        plain.load_state_dict(mine.state_dict())
        x1 = torch.randn(shape, generator=g).requires_grad_(True)
        x2 = x1.detach().clone().requires_grad_(True)
        PR.set_policy("3xtf32", split_backward=split_bwd)
        y1 = mine(x1)
        PR.set_policy("fp32")
        y2 = plain(x2)
        assert list(mine.state_dict()) == list(plain.state_dict())
        torch.testing.assert_close(y1, y2, rtol=1e-5, atol=2e-6)
        gy = torch.randn(y2.shape, generator=g)
        PR.set_policy("3xtf32")
        y1.backward(gy)
        y2.backward(gy)
        PR.set_policy("fp32", split_backward=False)
        torch.testing.assert_close(x1.grad, x2.grad, rtol=1e-5, atol=5e-6)
        torch.testing.assert_close(mine.weight.grad, plain.weight.grad, rtol=1e-5, atol=2e-5)
    hi = PR.tf32_round(torch.tensor([1.0 + 2 ** -11, 3.14159274]))
    assert (hi.view(torch.int32) & 0x1FFF).abs().sum() == 0


def test_bn_act_routes_cpu_tensors_through_torch_batchnorm():
    """norm.bn_act is what every conv block calls; without CUDA tensors (oracle comparisons, evaluation mode) it must be
    exactly relu(bn(x) + residual) of the stock module, running statistics included."""
    import torch.nn as nn
    from movedepth_b200 import norm as NM
    torch.manual_seed(0)
    x, r = torch.randn(3, 8, 5, 7), torch.randn(3, 8, 5, 7)
    a, b = nn.BatchNorm2d(8), nn.BatchNorm2d(8)
    want = torch.relu(a(x) + r)
    got = NM.bn_act(b, x, relu=True, residual=r)
    torch.testing.assert_close(got, want)
    torch.testing.assert_close(b.running_var, a.running_var)
    b.eval()
    a.eval()
    torch.testing.assert_close(NM.bn_act(b, x), a(x))


def test_workspace_and_buffer_size_queries_need_no_gpu():
    """Pure host entry points of the C ABI: wgrad workspaces (per-tile partials) and the peer-exchange buffer."""
    from movedepth_b200 import _lib
    from movedepth_b200.build import build
    build()                                   # no-op when the in-tree library is current
    L = _lib.lib()
    assert L.mvd_peer_allreduce_buffer_bytes(8, 2048) == 4096 + 2 * 8 * 2048 * 16      # header + 2 epoch slots x 8 ranks x 2048 entries of {lo, tag, hi, tag}
    assert L.mvd_peer_allreduce_buffer_bytes(0, 2048) == 0
    for fn, per_item in ((L.mvd_conv3d_c16o1_wgrad_workspace_bytes, 432 * 4), (L.mvd_conv3d_c16c16_wgrad_workspace_bytes, 6912 * 4),
                         (L.mvd_conv3d_c16c16_wgrad_tc_workspace_bytes, 2 * 6912 * 4)):
        n = fn(6, 96, 48, 160)
        assert n > 0 and n % per_item == 0 and n // per_item >= 6 * 6 * 5       # at least one item per (batch, tile)
        assert fn(0, 96, 48, 160) == 0


def test_custom_conv_modules_keep_reference_state_dict_and_cpu_path():
    """ProbConv3d / ConvBnReLUSeq are drop-ins: nn.Conv3d / nn.Sequential state-dict keys, stock arithmetic on CPU."""
    import torch.nn as nn
    from movedepth_b200 import networks as PN, precision as PR
    PR.set_policy("fp32")
    reg = PN.reg3d(16, 16, 3)
    keys = set(reg.state_dict().keys())
    assert "prob.weight" in keys and "conv7.0.weight" in keys and "conv7.1.running_mean" in keys and "conv0.bn.weight" in keys
    x = torch.randn(1, 16, 8, 8, 16)
    want = nn.functional.conv3d(x, reg.prob.weight, padding=1)
    torch.testing.assert_close(reg.prob(x), want)


def test_kitti_metric_stage_matches_the_oracle_and_the_eigen_crop():
    """movedepth_b200.evaluate_depth.kitti_metrics vs the oracle's restatement of evaluate_depth.py:259-331 (resize to the
    ground truth, Eigen crop, median scaling, clamp, metrics, oracle fusion), plus known answers: the Eigen crop of a
    375x1242 map is rows 153..370, columns 44..1196, and a perfect prediction scores zero error."""
    import numpy as np
    from movedepth_b200 import evaluate_depth as ED
    from oracle import evaluate as OE
    rng = np.random.default_rng(0)
    gt = [np.where(rng.random((375, 1242)) < 0.3, rng.uniform(0.5, 90, (375, 1242)), 0).astype(np.float32) for _ in range(2)]
    dz = rng.uniform(0.02, 1.0, (2, 48, 160)).astype(np.float32)
    dm = rng.uniform(0.02, 1.0, (2, 48, 160)).astype(np.float32)
    for split in ("eigen", "benchmark"):
        for scaling in (True, False):
            a, b = ED.kitti_metrics(dz, dm, gt, split, scaling), OE.kitti_metrics(dz, dm, gt, split, scaling)
            for k in ("mono", "mvs", "upbound"):
                np.testing.assert_allclose(a[k], b[k], rtol=1e-12)
            assert a["upbound"][0] <= min(a["mono"][0], a["mvs"][0]) + 1e-12        # the oracle fusion is never worse
    # only pixels inside the crop count: ground truth outside it may be anything
    g0 = np.full((375, 1242), 10.0, dtype=np.float32)
    g1 = g0.copy()
    g1[:153] = 55.0
    g1[371:] = 55.0
    g1[:, :44] = 55.0
    g1[:, 1197:] = 55.0
    d = np.full((1, 24, 80), 0.1, dtype=np.float32)
    r0, r1 = ED.kitti_metrics(d, d, [g0]), ED.kitti_metrics(d, d, [g1])
    np.testing.assert_allclose(r0["mvs"], r1["mvs"])
    np.testing.assert_allclose(r0["mvs"][:4], 0, atol=1e-6)                           # depth 10 everywhere == gt: zero errors
    g2 = g0.copy()
    g2[153, 44] = 55.0                                                                # first pixel inside the crop
    assert ED.kitti_metrics(d, d, [g2])["mvs"][0] > 0


def test_event_writer_tfrecord_and_protobuf_round_trip(tmp_path):
    """eventlog.SummaryWriter (replaces tensorboardX at movedepth/trainer.py:147-151, 772-793): CRC-32C known answer, TFRecord
    framing, scalar and image summaries read back, PNG payload decodes to the pixels written."""
    import struct
    import zlib
    from movedepth_b200 import eventlog
    assert eventlog.crc32c(b"123456789") == 0xE3069283                  # the Castagnoli check value
    assert eventlog.crc32c(b"") == 0
    w = eventlog.SummaryWriter(str(tmp_path))
    w.add_scalar("loss", 0.125, 7)
    w.add_scalar("abs_rel", 3.5, 2 ** 40)                               # a step that needs a multi-byte varint
    img = np.random.default_rng(0).random((3, 5, 9)).astype(np.float32)
    w.add_image("color_0_0/0", img, 8)
    w.add_image("disp_mono/0", eventlog.colormap(np.arange(12.0).reshape(3, 4)), 8)
    w.close()
    ev = list(eventlog.read_events(w.path))
    assert ev[0]["file_version"] == "brain.Event:2" and ev[0]["step"] == 0
    assert ev[1]["step"] == 7 and ev[1]["scalars"] == {"loss": 0.125}
    assert ev[2]["step"] == 2 ** 40 and ev[2]["scalars"] == {"abs_rel": 3.5}
    h, wd, c, png = ev[3]["images"]["color_0_0/0"]
    assert (h, wd, c) == (5, 9, 3) and png[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat = 8, b""
    while pos < len(png):                                               # chunk walk: length, type, data, CRC-32
        (n,) = struct.unpack(">I", png[pos:pos + 4])
        kind, body = png[pos + 4:pos + 8], png[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(kind + body)
        if kind == b"IDAT":
            idat += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(5, 1 + 9 * 3)
    assert (rows[:, 0] == 0).all()
    np.testing.assert_array_equal(rows[:, 1:].reshape(5, 9, 3), (np.moveaxis(img, 0, -1) * 255.0).astype(np.uint8))
    assert ev[4]["images"]["disp_mono/0"][:3] == (3, 4, 3)
    cm = eventlog.colormap(np.array([[0.0, 0.5, 1.0]]), normalize=False)
    np.testing.assert_allclose(cm[:, 0, 0] * 255, (13, 8, 135), atol=1e-4)
    np.testing.assert_allclose(cm[:, 0, 1] * 255, (204, 71, 120), atol=1e-4)
    np.testing.assert_allclose(cm[:, 0, 2] * 255, (240, 249, 33), atol=1e-4)
    with open(w.path, "r+b") as f:                                      # a flipped payload byte must be caught by the CRC
        f.seek(30)
        b = f.read(1)
        f.seek(30)
        f.write(bytes([b[0] ^ 1]))
    with pytest.raises(ValueError):
        list(eventlog.read_events(w.path))
