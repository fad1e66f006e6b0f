"""Inference path with the reference's semantics (movedepth/evaluate_depth.py:77-331) on the B200 kernels.

`DepthPredictor.predict(data)` is the body of the reference's inference loop (evaluate_depth.py:181-253): mono prior,
pose, FPN4 features, velocity-guided hypotheses (z-translation of batch item 0, line 218), the fused cost-volume kernel,
reg3d, fused softmax/local-max regression, convex upsampling, 1/depth.  `compute_errors` / `compute_fuse_errors` are the
KITTI metrics (22-64).  `evaluate(opt, dataloader, gt_depths)` strings them together for any loader that yields the
reference's item dicts; the KITTI datasets themselves are out of scope.
Checkpoints are the reference's files: `<folder>/<name>.pth`, loaded with strict=True like the reference does.
"""
import os

import numpy as np
import torch

from . import networks, ops
from .layers import convex_upsample_layer, disp_to_depth, fused_group_costvol, hypothesis_ratios, transformation_from_parameters
from . import precision as PR

MODEL_NAMES = ("mono_encoder", "mono_depth", "pose_encoder", "pose", "mvs_encoder", "reg3d", "up", "mask_cnn")


def build_models(opt):
    """The sub-models exactly as evaluate_depth.py:115-174 constructs them."""
    m = {}
    m["mono_encoder"] = networks.ResnetEncoder(opt.res_arch, False)
    m["mono_depth"] = networks.DepthDecoder(m["mono_encoder"].num_ch_enc)
    m["pose_encoder"] = networks.ResnetEncoder(opt.res_arch, False, num_input_images=2)
    m["pose"] = networks.PoseDecoder(m["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
    m["mvs_encoder"] = networks.FPN4(base_channels=8, scale=opt.prior_scale, dcn=opt.dcn)
    m["reg3d"] = networks.reg3d(opt.reg3d_c, opt.reg3d_c, 3) if opt.num_depth_bins >= 8 else networks.reg2d(opt.reg3d_c, opt.reg3d_c)
    m["up"] = convex_upsample_layer(8 * 2 ** opt.prior_scale, opt.prior_scale)
    m["mask_cnn"] = networks.UncertNet()
    for k in m:
        PR.adopt(m[k])
    return m


class DepthPredictor:
    def __init__(self, opt, models=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("movedepth_b200 has no CPU path: a CUDA device (B200, sm_100a) is required")
        ops._lib.lib()
        self.opt = opt
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.models = build_models(opt) if models is None else models
        PR.set_policy(getattr(opt, "b200_conv_precision", "3xtf32"))
        if models is None and getattr(opt, "load_weights_folder", None):
            self.load(opt.load_weights_folder)
        for m in self.models.values():
            m.to(self.device).eval()

    def load(self, folder):
        """strict=True per sub-model, like evaluate_depth.py:115-174 (`height`/`width` entries of mono_encoder.pth dropped)."""
        folder = os.path.expanduser(folder)
        for name in MODEL_NAMES:
            if name == "up" and not self.opt.convex_up:
                continue
            sd = torch.load(os.path.join(folder, "%s.pth" % name), map_location="cpu")
            sd = {k: v for k, v in sd.items() if k not in ("height", "width", "use_stereo")}
            self.models[name].load_state_dict(sd, strict=True)

    @torch.no_grad()
    def predict(self, data):
        """-> dict(pred_disp_z [B,H,W], pred_disp_mono [B,H,W], depth_mvs [B,H,W]) on the device."""
        o, m = self.opt, self.models
        cl = torch.channels_last
        data = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in data.items()}
        color = data[("color", 0, 0)].contiguous(memory_format=cl)
        out = m["mono_depth"](m["mono_encoder"](color))
        poses = []
        for f in o.frame_ids[1:]:
            other = data[("color", f, 0)]
            pair = [other, color] if f < 0 else [color, other]
            aa, tr = m["pose"]([m["pose_encoder"](torch.cat(pair, 1).contiguous(memory_format=cl))])
            poses.append(transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0)))
        rel = torch.stack(poses, 1)                                              # [B,S,4,4]
        ref_feat, ref_ctx = m["mvs_encoder"](color)
        src_feats = [m["mvs_encoder"](data[("color_aug", f, 0)].contiguous(memory_format=cl))[0] for f in o.matching_ids[1:]]
        B = color.shape[0]
        prior = 1.0 / (1.0 / o.max_depth + out[("disp", o.prior_scale)] * (1.0 / o.min_depth - 1.0 / o.max_depth))
        s = (o.depth_bin_fac * o.z_scale * rel[0, 0, 2, 3]).expand(B)            # batch item 0's z-translation for every item
        ratio = hypothesis_ratios(o.num_depth_bins, s, self.device, o.schedule_type)
        K, invK = data[("K", 2)], data[("inv_K", 2)]
        vols = [fused_group_costvol(ref_feat, src_feats[i], K, invK, rel[:, i], prior, ratio, o.reg3d_c, layout=ops.LAYOUT_BDHWG)
                for i in range(len(src_feats))]
        if len(vols) == 1:
            vol = vols[0]        # one view: the weight w/(1e-8+w) is 1 to 2e-7, whichever axis the softmax runs over
        else:
            wsum, acc = 1e-8, 0
            for v in vols:       # [B,G,D,h,w]: mean over G, softmax over D, max (evaluate_depth.py:236)
                wgt = torch.softmax(v.mean(1), dim=1).max(1)[0]
                wsum = wsum + wgt
                acc = acc + wgt[:, None, None] * v
            vol = acc / wsum[:, None, None]
        logits = m["reg3d"].forward_volume(vol)
        inv_a = 1.0 / (prior[:, 0] * ratio[:, -1].view(B, 1, 1))
        inv_b = 1.0 / (prior[:, 0] * ratio[:, 0].view(B, 1, 1))
        _, _, depth = ops.regress_depth(logits, inv_a, inv_b, o.norm_radius)
        if o.convex_up:
            depth = m["up"](depth, ref_ctx)
        scaled, _ = disp_to_depth(out[("disp", 0)], o.min_depth, o.max_depth)
        return dict(pred_disp_z=1.0 / depth, pred_disp_mono=scaled[:, 0], depth_mvs=depth)


def compute_errors(gt, pred):
    """abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 between ground-truth and predicted depths (evaluate_depth.py:22-40)."""
    ratio = np.maximum(gt / pred, pred / gt)
    a1, a2, a3 = [(ratio < 1.25 ** k).mean() for k in (1, 2, 3)]
    rmse = np.sqrt(((gt - pred) ** 2).mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    return np.mean(np.abs(gt - pred) / gt), np.mean((gt - pred) ** 2 / gt), rmse, rmse_log, a1, a2, a3


def compute_fuse_errors(gt, pred1, pred2):
    """Oracle fusion: per pixel the prediction closer to the ground truth (evaluate_depth.py:42-64)."""
    pick = np.abs(gt - pred1) < np.abs(pred2 - gt)
    return compute_errors(gt, np.where(pick, pred1, pred2))


def kitti_metrics(pred_disps_z, pred_disps_mono, gt_depths, eval_split="eigen", median_scaling=True, min_depth=1e-3, max_depth=80.0):
    """The metric stage of the reference's `evaluate` (movedepth/evaluate_depth.py:259-331): both disparity maps are resized
    to the ground-truth size (cv2.resize, bilinear), inverted, masked (`eigen`: 1e-3 < gt < 80 inside the Eigen crop;
    other splits: gt > 0), median-scaled per image unless `--disable_median_scaling`, clamped to [1e-3, 80] and scored
    with the seven KITTI metrics; `upbound` is the per-pixel oracle fusion.  Returns {"mono", "mvs", "upbound"} mean rows."""
    import cv2
    rows = {"mono": [], "mvs": [], "upbound": []}
    for i in range(len(pred_disps_mono)):
        gt = np.asarray(gt_depths[i])
        h, w = gt.shape[:2]
        disp_m = cv2.resize(np.squeeze(np.asarray(pred_disps_mono[i], dtype=np.float32)), (w, h))
        disp_z = cv2.resize(np.squeeze(np.asarray(pred_disps_z[i], dtype=np.float32)), (w, h))
        depth_z, depth_m = 1 / disp_z, 1 / disp_m
        if eval_split == "eigen":
            box = np.array([0.40810811 * h, 0.99189189 * h, 0.03594771 * w, 0.96405229 * w]).astype(np.int32)
            keep = np.zeros(gt.shape, dtype=bool)
            keep[box[0]:box[1], box[2]:box[3]] = True
            keep &= (gt > min_depth) & (gt < max_depth)
        else:
            keep = gt > 0
        depth_z, depth_m, g = depth_z[keep], depth_m[keep], gt[keep]
        if median_scaling:
            depth_m = depth_m * (np.median(g) / np.median(depth_m))
            depth_z = depth_z * (np.median(g) / np.median(depth_z))
        depth_z, depth_m = np.clip(depth_z, min_depth, max_depth), np.clip(depth_m, min_depth, max_depth)
        rows["mvs"].append(compute_errors(g, depth_z))
        rows["mono"].append(compute_errors(g, depth_m))
        rows["upbound"].append(compute_fuse_errors(g, depth_m, depth_z))
    return {k: np.array(v).mean(0) for k, v in rows.items()}


class GraphedPredictor:
    """`DepthPredictor.predict` captured in a CUDA graph for fixed-shape inference (batch-1 latency): inputs are copied into
    static device buffers, the whole forward (~330 launches) replays as one graph launch.  Note the reference takes the
    z-translation of batch item 0 for the whole batch (evaluate_depth.py:218); at batch 1 that is the item itself."""

    def __init__(self, predictor, example, warmup=2):
        self.pred = predictor
        dev = predictor.device
        self.static = {k: v.to(dev).clone() for k, v in example.items() if torch.is_tensor(v)}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # cuDNN autotuning and lazy initialisation outside the capture
                predictor.predict(self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.out = predictor.predict(self.static)

    def predict(self, data):
        for k, buf in self.static.items():
            buf.copy_(data[k], non_blocking=True)
        self.graph.replay()
        return self.out


def evaluate(opt, dataloader, gt_depths=None):
    """movedepth/evaluate_depth.py:77-331 for any loader that yields the reference's item dicts (the KITTI datasets
    themselves are out of scope): predictions with `DepthPredictor`, then `kitti_metrics` when ground truth is given
    (`opt.eval_split`, `opt.disable_median_scaling` honoured); otherwise the stacked disparities are returned."""
    pred = DepthPredictor(opt)
    dz, dm = [], []
    for data in dataloader:
        r = pred.predict(data)
        dz.append(r["pred_disp_z"].float().cpu().numpy())
        dm.append(r["pred_disp_mono"].float().cpu().numpy())
    dz, dm = np.concatenate(dz), np.concatenate(dm)
    if gt_depths is None:
        return dz, dm
    res = kitti_metrics(dz, dm, gt_depths, getattr(opt, "eval_split", "eigen"), not getattr(opt, "disable_median_scaling", False))
    for name in ("mono", "mvs", "upbound"):
        print("%s results:" % name)
        print(("{:>8} | " * 7).format("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3"))
        print(("&{: 8.3f}  " * 7).format(*res[name].tolist()) + "\\\\")
    return res
