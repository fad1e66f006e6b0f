"""Generate the golden vectors under tests/golden/ by running the REFERENCE code.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own `movedepth.layers`, `movedepth.networks` and `movedepth.trainer`
with the import shims of SURVEY.md §8(c) (stub tensorboardX/matplotlib/skimage/pykitti, PIL
ANTIALIAS alias, no-op torch.cuda.set_device, out-of-place residual in UncertNet.forward), gives
every sub-model the deterministic weights of tests/_weights.py, and records reference outputs
for (A) each operator on small seeded inputs and (B) whole `Trainer.process_batch` + backward
steps on a small configuration.  The inputs are regenerated from seeds by the tests
(tests/_cases.py), only outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("MOVEDEPTH_REFERENCE", "/root/reference")


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Writer:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, _):
            return lambda *a, **k: None

    stub("tensorboardX", SummaryWriter=_Writer)
    plt = stub("matplotlib.pyplot", get_cmap=lambda *a, **k: None)
    stub("matplotlib", pyplot=plt)
    stub("skimage.transform")
    stub("skimage", transform=sys.modules["skimage.transform"])
    stub("pykitti")
    from PIL import Image
    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    torch.cuda.set_device = lambda *_: None
    sys.path.insert(0, REF)
    import movedepth.layers as rl
    import movedepth.networks as rn
    import movedepth.trainer as rt
    from movedepth.options import MonodepthOptions

    def uncert_forward(self, x):       # same values as depth_decoder.py:387-393, residual out of place
        out = self.conv2(self.conv1(x))
        out = out + x
        return torch.sigmoid(self.head_convs(out))
    rn.UncertNet.forward = uncert_forward
    return rl, rn, rt, MonodepthOptions


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def op_vectors(rl):
    import _cases as C
    g = {}
    # A1: depth hypotheses
    c = C.case_hypotheses()
    g["hyp_v2"] = np32(rl.schedule_depth_rangev2(c["prior"], c["D"], c["fac"]))
    g["hyp_zv2"] = np32(rl.schedule_depth_range_zv2(c["prior"], c["D"], c["fac"], c["z_trans"]))
    g["hyp_v2_linear"] = np32(rl.schedule_depth_rangev2(c["prior"], c["D"], c["fac"], type="linear"))
    g["hyp_v2_log"] = np32(rl.schedule_depth_rangev2(c["prior"], c["D"], c["fac"], type="log"))
    g["hyp_zv2_log"] = np32(rl.schedule_depth_range_zv2(c["prior"], c["D"], c["fac"], c["z_trans"], type="log"))
    # A2: cost volume (reference layout [B,D,C,h,w]) for each pose case
    for name in C.COSTVOL_CASES:
        c = C.case_costvol(name)
        bp = rl.BackprojectDepth(c["D"], c["h"], c["w"])
        pj = rl.Project3D(c["D"], c["h"], c["w"])
        ref, src = c["ref"].clone().requires_grad_(True), c["src"].clone().requires_grad_(True)
        vol = rl.generate_costvol(ref, src, c["K"], c["invK"], c["hyps"], c["pose"], c["D"], bp, pj)
        g["costvol_%s" % name] = np32(vol)
        (vol * c["gvol"]).sum().backward()
        g["costvol_%s_gref" % name] = np32(ref.grad)
        g["costvol_%s_gsrc" % name] = np32(src.grad)
    # A3: entropy + localmax
    c = C.case_localmax()
    g["entropy"] = np32(rl.entropy(c["prob"], dim=1, keepdim=True))
    for r in (1, 2):
        g["localmax_r%d" % r] = np32(rl.localmax(c["prob"], r, c["D"], c["inv_a"], c["inv_b"]))
    g["localmax_onehot"] = np32(rl.localmax(c["onehot"], 1, c["D"], c["inv_a"], c["inv_b"]))
    # A4: convex upsample
    c = C.case_convex()
    g["convex_up"] = np32(rl.convex_upsample(c["depth"], c["mask"], 2))
    # A5: SSIM, reprojection pieces, smoothness
    c = C.case_images()
    g["ssim"] = np32(rl.SSIM()(c["x"], c["y"]))
    g["smooth"] = np32(rl.get_smooth_loss(c["disp"], c["x"]))
    # A6: pose matrices
    c = C.case_pose()
    g["T_fwd"] = np32(rl.transformation_from_parameters(c["aa"], c["tr"], invert=False))
    g["T_inv"] = np32(rl.transformation_from_parameters(c["aa"], c["tr"], invert=True))
    # A7: full-res border warp
    c = C.case_warp()
    bp = rl.BackprojectDepth(c["B"], c["H"], c["W"])
    pj = rl.Project3D(c["B"], c["H"], c["W"])
    grid = pj(bp(c["depth"], c["invK"]), c["K"], c["T"])
    g["warp_grid"] = np32(grid)
    g["warp_img"] = np32(torch.nn.functional.grid_sample(c["img"], grid, padding_mode="border", align_corners=True))
    g["reproj_loss"] = np32(0.85 * rl.SSIM()(torch.from_numpy(g["warp_img"]), c["img"].flip(0)).mean(1, True) +
                            0.15 * (c["img"].flip(0) - torch.from_numpy(g["warp_img"])).abs().mean(1, True))
    # A8: loss glue -- upsample + disp_to_depth (trainer.py:512-515), normalised smoothness (712-714), masked consistency (374, 398-400)
    import torch.nn.functional as F
    c = C.case_disp_pyramid()
    for s_ in range(4):
        up = F.interpolate(c["disp"][s_], [c["H"], c["W"]], mode="bilinear", align_corners=False)
        g["up_depth_s%d" % s_] = np32(rl.disp_to_depth(up, 0.1, 100.0)[1])
        d = c["disp"][s_]
        img = F.interpolate(c["img"], [c["H"] // 2 ** s_, c["W"] // 2 ** s_], mode="area") if s_ else c["img"]
        g["smooth_norm_s%d" % s_] = np32(rl.get_smooth_loss(d / (d.mean(2, True).mean(3, True) + 1e-7), img))
    c = C.case_masked()
    orig_ri = np.random.randint
    for i, xy in enumerate(c["boxes"]):
        q = list(xy)
        np.random.randint = lambda *a, **k: q.pop(0)
        try:
            _, m = rl.random_image_mask(torch.ones(c["B"], 3, c["H"], c["W"]), [c["H"] // 3, c["W"] // 3])
        finally:
            np.random.randint = orig_ri
        sel = F.interpolate(m, [c["h"], c["w"]], mode="bilinear", align_corners=True).sum(1).to(torch.bool)
        g["masked_sel_%d" % i] = sel.numpy()
        g["masked_loss_%d" % i] = np32(F.smooth_l1_loss(c["a"][sel], c["b"][sel], size_average=True) * 10 * 10)
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **g)
    print("ops.npz:", {k: v.shape for k, v in g.items()})


def step_vectors(rl, rn, rt, Options):
    import _cases as C
    from _weights import fill_deterministic
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--step=")]
    for name, cfg in C.STEP_CASES.items():
        if only and name not in only:
            continue
        argv = ["--no_cuda", "--weights_init", "scratch", "--num_workers", "0", "--data_path", "/nonexistent",
                "--png", "--log_dir", "/tmp/mvd_golden", "--prior_scale", "2", "--convex_up",
                "--height", str(cfg["H"]), "--width", str(cfg["W"]), "--num_depth_bins", str(cfg["D"]),
                "--batch_size", str(cfg["B"]), "--res_arch", str(cfg.get("arch", 18)), "--learning_rate", "2e-4",
                "--frame_ids"] + [str(f) for f in cfg["frame_ids"]]
        opt = Options().parser.parse_args(argv)
        tr = rt.Trainer(opt)
        for k, m in tr.models.items():
            fill_deterministic(m, salt=k + "/")
        tr.set_train()
        tr.epoch, tr.step = cfg["epoch"], 0
        inputs, noise, mask_xy = C.step_inputs(cfg)
        # the reference draws its own noise / box: feed it ours through the same RNG entry points
        noise_q = [n.clone() for n in noise]
        orig_randn = torch.randn
        torch.randn = lambda *a, **k: noise_q.pop(0) if (len(a) == 1 and isinstance(a[0], torch.Size)) else orig_randn(*a, **k)
        xy_q = list(mask_xy)
        orig_ri = np.random.randint
        np.random.randint = lambda *a, **k: xy_q.pop(0)
        captured = {}
        def grab(_m, args):                     # first call = un-augmented reference frame
            captured.setdefault("vol", args[0].detach().clone())
        h1 = tr.models["reg3d"].register_forward_pre_hook(grab)
        try:
            outputs, losses = tr.process_batch(dict(inputs), is_train=True)
        finally:
            torch.randn, np.random.randint = orig_randn, orig_ri
            h1.remove()
        tr.model_optimizer.zero_grad()
        losses["loss"].backward()
        g = {"loss/" + k: np32(v) for k, v in losses.items()}
        for s in range(4):
            g["disp%d" % s] = np32(outputs[("disp", s)])
        g["depth_mvs"] = np32(outputs["depth_mvs"])
        g["masked_depth"] = np32(outputs["masked_depth"])
        g["fused_depth"] = np32(outputs["fused_depth"])
        g["trust_mono_mask"] = np32(outputs["trust_mono_mask"])
        g["mono_reproj_loss"] = np32(outputs["mono_reproj_loss"])
        g["mvs_reprojection_loss"] = np32(outputs["mvs_reprojection_loss"])
        if not cfg.get("slim"):                 # full-size cases keep the maps the depth bar is stated on, not the volume
            g["cost_volume"] = np32(captured["vol"])
        else:
            for k in ("mono_reproj_loss", "mvs_reprojection_loss", "masked_depth"):
                g.pop(k)
        for f in cfg["frame_ids"][1:]:
            g["cam_T_cam_%d" % f] = np32(outputs[("cam_T_cam", 0, f)])
            if not cfg.get("slim"):
                g["warped_%d_s0" % f] = np32(outputs[("color", f, 0)])
        for k, m in tr.models.items():
            sq = sum(float((p.grad.double() ** 2).sum()) for p in m.parameters() if p.grad is not None)
            g["gradnorm/" + k] = np.float64(sq ** 0.5)
        sd = {k: dict(m.named_parameters()) for k, m in tr.models.items()}
        for mk, pk in C.GRAD_PROBES:
            g["grad/%s/%s" % (mk, pk)] = np32(sd[mk][pk].grad)
        # one Adam step, then probe a few updated parameters
        tr.model_optimizer.step()
        for mk, pk in C.GRAD_PROBES:
            g["adam/%s/%s" % (mk, pk)] = np32(sd[mk][pk])
        np.savez_compressed(os.path.join(HERE, "step_%s.npz" % name), **g)
        print("step_%s.npz: loss=%.6f" % (name, float(losses["loss"])),
              {k: float(v) for k, v in g.items() if k.startswith("loss/")})


def eval_vectors(rl, rn):
    """The inference loop body of movedepth/evaluate_depth.py:181-253 re-assembled from the REFERENCE's own functions
    and networks (the loop itself is inline in evaluate(), which needs the KITTI dataset), eval mode, deterministic
    weights, one smooth synthetic batch."""
    import _cases as C
    from _weights import fill_deterministic
    import torch.nn.functional as F
    cfg = C.EVAL_CASE
    opt = C.step_options(cfg)
    m = {}
    m["mono_encoder"] = rn.ResnetEncoder(num_layers=18, pretrained=False)
    m["mono_depth"] = rn.DepthDecoder(m["mono_encoder"].num_ch_enc, match_conv=False, ddv=False, discret=None, mono_conf=False, mono_bins=None)
    m["pose_encoder"] = rn.ResnetEncoder(18, False, num_input_images=2)
    m["pose"] = rn.PoseDecoder(m["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
    m["mvs_encoder"] = rn.FPN4(base_channels=8, scale=opt.prior_scale, dcn=False)
    m["reg3d"] = rn.reg3d(in_channels=opt.reg3d_c, base_channels=opt.reg3d_c, down_size=3)
    m["up"] = rl.convex_upsample_layer(feature_dim=8 * 2 ** opt.prior_scale, scale=opt.prior_scale)
    for k, mod in m.items():
        fill_deterministic(mod, salt=k + "/")
        mod.eval()
    data, _, _ = C.step_inputs(cfg)
    h, w = cfg["H"] // 4, cfg["W"] // 4
    bp, pj = rl.BackprojectDepth(cfg["D"], h, w), rl.Project3D(cfg["D"], h, w)
    with torch.no_grad():
        color = data["color", 0, 0]
        out = m["mono_depth"](m["mono_encoder"](color))
        poses = []
        for f in cfg["frame_ids"][1:]:
            pair = [data["color", f, 0], color] if f < 0 else [color, data["color", f, 0]]
            aa, tr = m["pose"]([m["pose_encoder"](torch.cat(pair, 1))])
            poses.append(rl.transformation_from_parameters(aa[:, 0], tr[:, 0], invert=f < 0))
        rel = torch.stack(poses, 1)
        ref_feat, ref_ctx = m["mvs_encoder"](color)
        src_feat, _ = m["mvs_encoder"](data["color_aug", -1, 0])
        disp_prior = out[("disp", opt.prior_scale)]
        depth_prior = 1 / (1 / opt.max_depth + disp_prior * (1 / opt.min_depth - 1 / opt.max_depth))
        z_scale = opt.z_scale * rel[0, 0, 2, -1]
        hyps = rl.schedule_depth_range_zv2(depth_prior, ndepth=cfg["D"], scale_fac=opt.depth_bin_fac, z_trans=z_scale)
        cv = rl.generate_costvol(ref_feat, src_feat, data["K", 2], data["inv_K", 2], hyps, rel[:, 0:1], cfg["D"], bp, pj)
        B, D, Cc, H, W = cv.shape
        cv = cv.reshape(B, D, -1, opt.reg3d_c, H, W).mean(2)
        wgt = torch.softmax(cv.mean(2), dim=1).max(1)[0]
        feats = (wgt.unsqueeze(1).unsqueeze(1) * cv) / (1e-8 + wgt).unsqueeze(1).unsqueeze(1)
        prob = F.softmax(m["reg3d"](feats), 1)
        depth = rl.localmax(prob, opt.norm_radius, cfg["D"], 1 / hyps[:, -1], 1 / hyps[:, 0])
        depth_up = m["up"](depth, ref_ctx)
        disp_mono, _ = rl.disp_to_depth(out[("disp", 0)], opt.min_depth, opt.max_depth)
    g = dict(pred_disp_z=np32(1 / depth_up), pred_disp_mono=np32(disp_mono[:, 0]), depth_lowres=np32(depth), tz0=np32(rel[0, 0, 2, -1]))
    np.savez_compressed(os.path.join(HERE, "eval_r18.npz"), **g)
    print("eval_r18.npz:", {k: v.shape for k, v in g.items()}, "tz0", float(g["tz0"]))


def datapipe_vectors():
    """The reference's own MonoDataset.preprocess (movedepth/datasets/mono_dataset.py:104-126) on two synthetic frames:
    pyramid of PIL LANCZOS resizes + ToTensor, with and without the horizontal flip of `get_color` (kitti_dataset.py) and
    with `transforms.ColorJitter` seeded through torch's global RNG."""
    import _cases as C
    import movedepth.datasets as ds
    from PIL import Image
    from torchvision import transforms
    c = C.case_frames()
    d = ds.KITTIRAWDataset("/nonexistent", ["2011_09_26/x 0 l"], c["H"], c["W"], [0, -1], 4, is_train=True, img_ext=".png")
    g = {}
    for tag, flip, seed in (("plain", False, None), ("flip", True, None), ("jitter", False, c["seed"])):
        inputs = {}
        for f in (0, -1):
            im = Image.fromarray(c["frames"][f], "RGB")
            inputs[("color", f, -1)] = im.transpose(Image.FLIP_LEFT_RIGHT) if flip else im
        if seed is None:
            aug = (lambda x: x)
        else:
            torch.manual_seed(seed)
            aug = transforms.ColorJitter(d.brightness, d.contrast, d.saturation, d.hue)
        d.preprocess(inputs, aug)
        for f in (0, -1):
            for s_ in range(4):
                g["%s_color_%d_%d" % (tag, f, s_)] = np32(inputs[("color", f, s_)])
                if seed is not None:
                    g["%s_color_aug_%d_%d" % (tag, f, s_)] = np32(inputs[("color_aug", f, s_)])
    np.savez_compressed(os.path.join(HERE, "datapipe.npz"), **g)
    print("datapipe.npz:", len(g), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    rl, rn, rt, Options = import_reference()
    if "--eval-only" not in sys.argv and "--datapipe-only" not in sys.argv:
        if "--steps-only" not in sys.argv:
            op_vectors(rl)
        if "--ops-only" not in sys.argv:
            step_vectors(rl, rn, rt, Options)
    if "--datapipe-only" in sys.argv:
        datapipe_vectors()
        sys.exit(0)
    if "--ops-only" not in sys.argv and "--steps-only" not in sys.argv:
        eval_vectors(rl, rn)
        datapipe_vectors()
