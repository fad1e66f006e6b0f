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
