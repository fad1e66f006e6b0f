"""Oracle restatement of `Trainer.process_batch` + the optimiser step (CPU, plain torch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows movedepth/trainer.py:297-442
(forward + losses), 445-468 (poses), 491-532 (image warps), 535-724 (losses) and 265-272
(zero_grad / backward / Adam.step).  It is also what `bench.py --impl reference` and the
`cpu_baseline` leg time on the host cores (kind "port": the Python reference itself cannot
travel to the GPU box).
"""
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import layers as L
from . import networks as N


def default_options(**kw):
    """The hot-path flags and their reference defaults (movedepth/options.py:7-350), with the
    canonical training overrides of train_movedepth.sh:16-30 (`--prior_scale 2 --convex_up`)."""
    o = dict(height=192, width=640, scales=[0, 1, 2, 3], frame_ids=[0, -1], matching_ids=[0, -1],
             num_depth_bins=16, reg3d_c=16, prior_scale=2, depth_bin_fac=0.3, z_scale=30.0,
             ztrans_start_epc=8, schedule_type="inverse", norm_radius=1, min_depth=0.1,
             max_depth=100.0, ssim_lw=0.85, no_ssim=False, disable_automasking=False,
             disparity_smoothness=1e-3, mask_lw=10.0, convex_up=True, mask_mvs_auto=False,
             mvs_smooth_loss=False, res_arch=18, batch_size=2, learning_rate=2e-4, lr_fac=1.0)
    o.update(kw)
    return SimpleNamespace(**o)


def build_models(opt):
    """movedepth/trainer.py:65-131 -- the 8 sub-models, keyed as `Trainer.models`."""
    m = {}
    m["mono_encoder"] = N.ResnetEncoder(opt.res_arch, False)
    m["mono_depth"] = N.DepthDecoder(m["mono_encoder"].num_ch_enc, opt.scales)
    m["pose_encoder"] = N.ResnetEncoder(opt.res_arch, False, num_input_images=2)
    m["pose"] = N.PoseDecoder(m["pose_encoder"].num_ch_enc, 1, 2)
    m["mask_cnn"] = N.UncertNet()
    m["mvs_encoder"] = N.FPN4(base_channels=8, scale=opt.prior_scale)
    m["reg3d"] = N.Reg3d(opt.reg3d_c, opt.reg3d_c, 3)
    m["up"] = N.ConvexUpsampleLayer(8 * 2 ** opt.prior_scale, opt.prior_scale)
    return m


def param_groups(models, opt):
    """movedepth/trainer.py:62-141: group 0 (lr) = mono_encoder, mono_depth, pose_encoder, pose, up;
    group 1 (lr*lr_fac) = mask_cnn, mvs_encoder, reg3d."""
    g0 = [p for k in ("mono_encoder", "mono_depth", "pose_encoder", "pose", "up") for p in models[k].parameters()]
    g1 = [p for k in ("mask_cnn", "mvs_encoder", "reg3d") for p in models[k].parameters()]
    return [{"params": g0, "lr": opt.learning_rate}, {"params": g1, "lr": opt.learning_rate * opt.lr_fac}]


class OracleStep:
    def __init__(self, opt, models=None):
        self.opt = opt
        self.models = models if models is not None else build_models(opt)
        self.optimizer = torch.optim.Adam(param_groups(self.models, opt))
        for m in self.models.values():
            m.train()

    # ---- poses (trainer.py:445-468)
    def predict_poses(self, inputs, out):
        o = self.opt
        for f in o.frame_ids[1:]:
            pair = [inputs["color_aug", f, 0], inputs["color_aug", 0, 0]] if f < 0 else \
                   [inputs["color_aug", 0, 0], inputs["color_aug", f, 0]]
            feats = [self.models["pose_encoder"](torch.cat(pair, 1))]
            aa, tr = self.models["pose"](feats)
            out["axisangle", 0, f], out["translation", 0, f] = aa, tr
            out["cam_T_cam", 0, f] = L.transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0))
        for f in o.matching_ids[1:]:
            inputs["relative_pose", f] = out["cam_T_cam", 0, f].clone().detach()

    # ---- mono photometric loss (trainer.py:510-532 + 675-724)
    def mono_losses(self, inputs, out, noise):
        o = self.opt
        losses, total = {}, 0
        target = inputs["color", 0, 0]
        for s in o.scales:
            disp = out["disp", s]
            disp_full = F.interpolate(disp, [o.height, o.width], mode="bilinear", align_corners=False)
            _, depth = L.disp_to_depth(disp_full, o.min_depth, o.max_depth)
            out["depth", 0, s] = depth
            reproj = []
            for f in o.frame_ids[1:]:
                pred, grid = L.warp_image(inputs["color", f, 0], depth, inputs["K", 0], inputs["inv_K", 0],
                                          out["cam_T_cam", 0, f])
                out["sample", f, s], out["color", f, s] = grid, pred
                reproj.append(L.reprojection_loss(pred, target, o.ssim_lw, o.no_ssim))
            reproj = torch.cat(reproj, 1).min(1, keepdim=True)[0]
            if not o.disable_automasking:
                ident = torch.cat([L.reprojection_loss(inputs["color", f, 0], target, o.ssim_lw, o.no_ssim)
                                   for f in o.frame_ids[1:]], 1).min(1, keepdim=True)[0]
                ident = ident + noise.pop(0) * 1e-5
                mask = (torch.argmin(torch.cat([reproj, ident], 1), 1, keepdim=True) == 0).float()
            else:
                mask = torch.ones_like(reproj)
            if s == 0:
                out["mono_reproj_loss"] = reproj
            loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
            norm_disp = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
            sm = L.smooth_loss(norm_disp, inputs["color", 0, s])
            losses["mono_smooth_loss/{}".format(s)] = sm
            loss = loss + o.disparity_smoothness * sm / (2 ** s)
            total = total + loss
            losses["loss/{}".format(s)] = loss
        losses["loss"] = total / len(o.scales)
        return losses

    # ---- cost volume -> reg3d -> probabilities (trainer.py:349-367)
    def regress(self, ref_feat, src_feats, inputs, hyps, poses):
        o = self.opt
        vols = []
        for i in range(len(src_feats)):
            cv = L.cost_volume(ref_feat, src_feats[i], inputs["K", 2], inputs["inv_K", 2], hyps, poses[:, i:i + 1])
            vols.append(L.group_correlation(cv, o.reg3d_c))
        feats = L.fuse_views(vols)
        logits = self.models["reg3d"](feats)
        return F.softmax(logits, 1), feats

    def process_batch(self, inputs, epoch=0, noise=None, mask_xy=None):
        """inputs: dict per movedepth/datasets/mono_dataset.py:134-154.  `noise`: optional list of
        [B,1,H,W] N(0,1) tensors consumed in the order the reference draws them (one per mono
        scale; then fuse, then mvs when mask_mvs_auto); drawn from torch's CPU generator if None.
        `mask_xy`: optional (x, y) of the augmentation box, else np.random like the reference."""
        o = self.opt
        B = inputs["color_aug", 0, 0].shape[0]
        if noise is None:
            n_noise = (0 if o.disable_automasking else len(o.scales)) + (2 if o.mask_mvs_auto else 0)
            noise = [torch.randn(B, 1, o.height, o.width) for _ in range(n_noise)]
        noise = list(noise)
        out = {}
        self.predict_poses(inputs, out)
        poses = torch.stack([inputs["relative_pose", f] for f in o.matching_ids[1:]], 1)   # [B,M,4,4]

        ref_feat, ref_ctx = self.models["mvs_encoder"](inputs["color_aug", 0, 0])
        src_feats = [self.models["mvs_encoder"](inputs["color_aug", f, 0])[0] for f in o.matching_ids[1:]]

        out.update(self.models["mono_depth"](self.models["mono_encoder"](inputs["color_aug", 0, 0])))
        losses = self.mono_losses(inputs, out, noise)

        # mono prior -> hypotheses (trainer.py:333-346)
        disp_prior = out["disp", o.prior_scale].clone().detach()
        depth_prior = 1 / (1 / o.max_depth + disp_prior * (1 / o.min_depth - 1 / o.max_depth))
        if epoch > o.ztrans_start_epc:
            hyps = L.depth_hypotheses(depth_prior, o.num_depth_bins, o.depth_bin_fac,
                                      z_trans=o.z_scale * poses[:, :, 2:3, -1:], kind=o.schedule_type)
        else:
            hyps = L.depth_hypotheses(depth_prior, o.num_depth_bins, o.depth_bin_fac, kind=o.schedule_type)
        out["depth_hypotheses"] = hyps

        prob, feats = self.regress(ref_feat, src_feats, inputs, hyps, poses)
        out["cost_volume"], out["cost_prob"] = feats, prob
        ent = L.entropy(prob, dim=1, keepdim=True)
        trust = self.models["mask_cnn"](ent)
        inv_a, inv_b = 1 / hyps[:, -1], 1 / hyps[:, 0]
        depth_mvs = L.localmax(prob, o.norm_radius, o.num_depth_bins, inv_a, inv_b)
        out["depth_mvs_lowres"] = depth_mvs

        # masked-augmentation consistency (trainer.py:374-403)
        masked_img, aug_mask = L.box_mask(inputs["color_aug", 0, 0], [o.height // 3, o.width // 3], mask_xy)
        aug_feat, _ = self.models["mvs_encoder"](masked_img)
        prob_aug, _ = self.regress(aug_feat, src_feats, inputs, hyps, poses)
        depth_aug = L.localmax(prob_aug, o.norm_radius, o.num_depth_bins, inv_a, inv_b)
        sel = F.interpolate(aug_mask, list(depth_aug.shape[1:]), mode="bilinear", align_corners=True).sum(1).to(torch.bool)
        masked = F.smooth_l1_loss(depth_aug[sel], depth_mvs[sel], reduction="mean") * o.mask_lw
        losses["masked_loss"] = masked * o.mask_lw          # weight applied twice (reference quirk)
        losses["loss"] = losses["loss"] + losses["masked_loss"]
        out["masked_depth"], out["masked_aug"] = depth_aug, aug_mask

        # upsample + fuse (trainer.py:405-416)
        if o.convex_up:
            depth_up = self.models["up"](depth_mvs, ref_ctx)
        else:
            depth_up = F.interpolate(depth_mvs.unsqueeze(1), [o.height, o.width], mode="bilinear", align_corners=True)[:, 0]
        out["depth_mvs"] = depth_up
        _, mono_depth = L.disp_to_depth(out["disp", 0], o.min_depth, o.max_depth)
        trust = F.interpolate(trust, [o.height, o.width], mode="bilinear", align_corners=True)
        fused = (1 - trust) * depth_up[:, None].detach() + trust * mono_depth.detach()
        out["fused_depth"], out["trust_mono_mask"] = fused, trust

        # fuse loss: L1 only (ssim_lw=0), poses detached (trainer.py:569-612)
        target = inputs["color", 0, 0]
        rl = []
        for f in o.frame_ids[1:]:
            pred, _ = L.warp_image(inputs["color", f, 0], fused, inputs["K", 0], inputs["inv_K", 0],
                                   out["cam_T_cam", 0, f].detach())
            out["mvs_color_fuse", f] = pred
            rl.append(L.reprojection_loss(pred, target, 0, o.no_ssim))
        rl = torch.cat(rl, 1).min(1, keepdim=True)[0]
        if o.mask_mvs_auto:
            ident = torch.cat([L.reprojection_loss(inputs["color", f, 0], target, 0, o.no_ssim)
                               for f in o.frame_ids[1:]], 1).min(1, keepdim=True)[0]
            ident = ident + noise.pop(0) * 1e-5
            fmask = (torch.argmin(torch.cat([rl, ident], 1), 1, keepdim=True) == 0).float()
        else:
            fmask = torch.ones_like(rl)
        fuse_loss = (rl * fmask).sum() / (fmask.sum() + 1e-7)

        # mvs loss: SSIM+L1 at scale 0 with depth_mvs, poses detached, mask = ones (trainer.py:495-508, 621-673)
        rl = []
        for f in o.frame_ids[1:]:
            pred, _ = L.warp_image(inputs["color", f, 0], depth_up, inputs["K", 0], inputs["inv_K", 0],
                                   out["cam_T_cam", 0, f].detach())
            out["mvs_color", f] = pred
            rl.append(L.reprojection_loss(pred, target, o.ssim_lw, o.no_ssim))
        rl = torch.cat(rl, 1).min(1, keepdim=True)[0]
        if o.mask_mvs_auto:
            noise.pop(0)                                   # drawn by the reference, mask then overwritten by ones
        out["mvs_reprojection_loss"] = rl
        mvs_loss = rl.sum() / (rl.numel() + 1e-7)
        if o.mvs_smooth_loss:
            d = depth_up.unsqueeze(1)
            sm = L.smooth_loss(d / (d.mean(2, True).mean(3, True) + 1e-7), inputs["color", 0, 0])
            losses["mvs_smooth_loss/0"] = sm
            mvs_loss = mvs_loss + o.disparity_smoothness * sm
        out["mvs_reproj_loss"] = mvs_loss

        # merge (trainer.py:429-440): loss = mvs + (mono + masked) + fuse
        losses["fuse_reproj_loss"] = fuse_loss
        losses["loss"] = mvs_loss + losses["loss"] + fuse_loss
        return out, losses

    def train_step(self, inputs, epoch=0, noise=None, mask_xy=None):
        """movedepth/trainer.py:269-272."""
        out, losses = self.process_batch(inputs, epoch, noise, mask_xy)
        self.optimizer.zero_grad()
        losses["loss"].backward()
        self.optimizer.step()
        return out, losses


def synthetic_inputs(opt, batch, seed=1, smooth=False, shift_px=2):
    """Synthetic KITTI-shape item dict (schema movedepth/datasets/mono_dataset.py:134-154,
    intrinsics kitti_dataset.py:26-29 scaled per mono_dataset.py:209-218).  `smooth=True` uses bicubic-
    upsampled low-res noise with the source frames shifted by `shift_px` (well-conditioned parity case)."""
    g = torch.Generator().manual_seed(seed)
    H, W = opt.height, opt.width
    inputs = {}
    base = None
    for f in opt.frame_ids:
        if smooth:
            if base is None:
                lo = torch.rand(batch, 3, H // 8, (W + 64) // 8, generator=g)
                base = F.interpolate(lo, size=(H, W + 64), mode="bicubic", align_corners=False).clamp(0, 1)
            off = 32 + shift_px * f
            img = base[:, :, :, off:off + W].contiguous()
        else:
            img = torch.rand(batch, 3, H, W, generator=g)
        for s in range(4):
            im = img if s == 0 else F.interpolate(img, size=(H // 2 ** s, W // 2 ** s), mode="area")
            inputs["color", f, s] = im
            inputs["color_aug", f, s] = im.clone()
    for s in range(4):
        K = torch.tensor([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float64)
        K[0] *= W // 2 ** s
        K[1] *= H // 2 ** s
        inputs["K", s] = K.float().repeat(batch, 1, 1)
        inputs["inv_K", s] = torch.linalg.pinv(K).float().repeat(batch, 1, 1)
    return inputs
