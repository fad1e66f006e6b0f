"""Oracle restatement of the reference's inference path (CPU, plain torch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows movedepth/evaluate_depth.py:181-253: mono encoder /
decoder, pose net per source frame, FPN4 matching features, velocity-guided hypotheses around the mono prior with the
z-translation of batch item 0 (line 218 -- a quirk that is part of the behaviour), cost volume -> group mean -> view
weight (softmax over the DEPTH axis there, line 236) -> reg3d -> softmax -> localmax -> convex upsampling -> 1/depth.
The metric code follows evaluate_depth.py:22-40.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import layers as L


@torch.no_grad()
def predict(models, data, opt):
    """models: dict keyed like Trainer.models, in eval mode; data: item dict (mono_dataset.py:134-154) with frames
    opt.frame_ids.  Returns dict(pred_disp_z [B,H,W], pred_disp_mono [B,H,W], depth_mvs [B,H,W])."""
    color = data["color", 0, 0]
    out = models["mono_depth"](models["mono_encoder"](color))
    poses = []
    for f in opt.frame_ids[1:]:
        pair = [data["color", f, 0], color] if f < 0 else [color, data["color", f, 0]]
        aa, tr = models["pose"]([models["pose_encoder"](torch.cat(pair, 1))])
        poses.append(L.transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0)))
    rel = torch.stack(poses, 1)                                                  # [B,S,4,4]
    ref_feat, ref_ctx = models["mvs_encoder"](color)
    src_feats = [models["mvs_encoder"](data["color_aug", f, 0])[0] for f in opt.matching_ids[1:]]
    disp_prior = out["disp", opt.prior_scale]
    depth_prior = 1 / (1 / opt.max_depth + disp_prior * (1 / opt.min_depth - 1 / opt.max_depth))
    z_scale = opt.z_scale * rel[0, 0, 2, -1]                                     # batch item 0 only (evaluate_depth.py:218)
    hyps = L.depth_hypotheses(depth_prior, opt.num_depth_bins, opt.depth_bin_fac, z_trans=z_scale, kind=opt.schedule_type)
    vols = []
    for i in range(len(src_feats)):
        cv = L.cost_volume(ref_feat, src_feats[i], data["K", 2], data["inv_K", 2], hyps, rel[:, i:i + 1])
        vols.append(L.group_correlation(cv, opt.reg3d_c))
    feats = L.fuse_views(vols, eval_axis=True)
    prob = F.softmax(models["reg3d"](feats), 1)
    depth = L.localmax(prob, opt.norm_radius, opt.num_depth_bins, 1 / hyps[:, -1], 1 / hyps[:, 0])
    if opt.convex_up:
        depth = models["up"](depth, ref_ctx)
    scaled, _ = L.disp_to_depth(out["disp", 0], opt.min_depth, opt.max_depth)
    return dict(pred_disp_z=1 / depth, pred_disp_mono=scaled[:, 0], depth_mvs=depth)


def compute_errors(gt, pred):
    """abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 (movedepth/evaluate_depth.py:22-40)."""
    ratio = np.maximum(gt / pred, pred / gt)
    a1, a2, a3 = [(ratio < 1.25 ** k).mean() for k in (1, 2, 3)]
    rmse = np.sqrt(((gt - pred) ** 2).mean())
    rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    return np.mean(np.abs(gt - pred) / gt), np.mean((gt - pred) ** 2 / gt), rmse, rmse_log, a1, a2, a3


def compute_fuse_errors(gt, pred1, pred2):
    """movedepth/evaluate_depth.py:42-64 -- per pixel the prediction closer to the ground truth ("upbound" row)."""
    pick = np.abs(gt - pred1) < np.abs(pred2 - gt)
    return compute_errors(gt, np.where(pick, pred1, pred2))


def kitti_metrics(pred_disps_z, pred_disps_mono, gt_depths, eval_split="eigen", median_scaling=True, min_depth=1e-3, max_depth=80.0):
    """movedepth/evaluate_depth.py:259-331 -- per image: cv2.resize of both disparities to the ground-truth size, 1/disp,
    Eigen crop (0.408..0.992 of the height, 0.036..0.964 of the width; int32 truncation) AND 1e-3 < gt < 80 mask for the
    `eigen` split (gt > 0 otherwise), per-image median scaling unless disabled, clamp to [1e-3, 80], the seven metrics for
    the multi-frame, the mono and the oracle-fused prediction; returns the three mean rows (mono, mvs, upbound)."""
    import cv2
    rows = {"mono": [], "mvs": [], "upbound": []}
    for i in range(pred_disps_mono.shape[0]):
        gt = gt_depths[i]
        gh, gw = gt.shape[:2]
        dm = cv2.resize(np.squeeze(pred_disps_mono[i]), (gw, gh))
        dz = cv2.resize(np.squeeze(pred_disps_z[i]), (gw, gh))
        pz, pm = 1 / dz, 1 / dm
        if eval_split == "eigen":
            mask = np.logical_and(gt > min_depth, gt < max_depth)
            crop = np.array([0.40810811 * gh, 0.99189189 * gh, 0.03594771 * gw, 0.96405229 * gw]).astype(np.int32)
            cm = np.zeros(mask.shape)
            cm[crop[0]:crop[1], crop[2]:crop[3]] = 1
            mask = np.logical_and(mask, cm)
        else:
            mask = gt > 0
        pz, pm, g = pz[mask], pm[mask], gt[mask]
        if median_scaling:
            pm = pm * (np.median(g) / np.median(pm))
            pz = pz * (np.median(g) / np.median(pz))
        pz = np.clip(pz, min_depth, max_depth)
        pm = np.clip(pm, min_depth, max_depth)
        rows["mvs"].append(compute_errors(g, pz))
        rows["mono"].append(compute_errors(g, pm))
        rows["upbound"].append(compute_fuse_errors(g, pm, pz))
    return {k: np.array(v).mean(0) for k, v in rows.items()}
