"""Oracle restatement of the geometry / cost-volume / loss operators.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference lines (relative to /root/reference/) whose behaviour it restates.  The
quirks listed in SURVEY.md §8(c) are reproduced on purpose, not "fixed".
"""
import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# depth <-> disparity, poses
# ----------------------------------------------------------------------------
def disp_to_depth(disp, min_depth, max_depth):
    """movedepth/layers.py:400-409 -- sigmoid output -> (scaled disparity, depth)."""
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1.0 / scaled


def rotation_from_axisangle(vec):
    """movedepth/layers.py:479-518 -- Rodrigues formula, input [B,1,3] -> [B,4,4].

    The `+1e-7` on the angle before normalising the axis is part of the contract.
    """
    angle = vec.norm(p=2, dim=2, keepdim=True)                 # [B,1,1]
    axis = vec / (angle + 1e-7)
    c, s = torch.cos(angle), torch.sin(angle)
    omc = 1 - c
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))    # each [B,1,1]
    out = vec.new_zeros(vec.shape[0], 4, 4)
    sq = torch.squeeze
    # multiplication order below follows the reference ((a*C) first, then *b) so fp32 rounding agrees
    xC, yC, zC = x * omc, y * omc, z * omc
    out[:, 0, 0] = sq(x * xC + c)
    out[:, 0, 1] = sq(x * yC - z * s)
    out[:, 0, 2] = sq(z * xC + y * s)
    out[:, 1, 0] = sq(x * yC + z * s)
    out[:, 1, 1] = sq(y * yC + c)
    out[:, 1, 2] = sq(y * zC - x * s)
    out[:, 2, 0] = sq(z * xC - y * s)
    out[:, 2, 1] = sq(y * zC + x * s)
    out[:, 2, 2] = sq(z * zC + c)
    out[:, 3, 3] = 1
    return out


def translation_matrix(t):
    """movedepth/layers.py:464-477 -- [B,1,3] (or [B,3]) -> homogeneous [B,4,4]."""
    n = t.shape[0]
    out = torch.eye(4, dtype=t.dtype, device=t.device).repeat(n, 1, 1)
    out[:, :3, 3] = t.reshape(n, 3)
    return out


def transformation_from_parameters(axisangle, translation, invert=False):
    """movedepth/layers.py:412-429 -- pose-net output -> 4x4.  invert => R^T and -t, M = R.T"""
    rot = rotation_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        rot = rot.transpose(1, 2)
        t = t * -1
    trans = translation_matrix(t)
    return torch.matmul(rot, trans) if invert else torch.matmul(trans, rot)


# ----------------------------------------------------------------------------
# pinhole geometry
# ----------------------------------------------------------------------------
def pixel_grid(h, w, dtype=torch.float32, device="cpu"):
    """Homogeneous pixel coordinates [3, h*w], x fastest (movedepth/layers.py:566-579)."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=dtype, device=device),
                            torch.arange(w, dtype=dtype, device=device), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, dtype=dtype, device=device)], 0)


def backproject(depth, inv_K, h, w):
    """movedepth/layers.py:581-586 -- depth [N,1,h,w] (any shape with N*h*w elems), inv_K [N or 1,4,4]
    -> camera points [N,4,h*w] (homogeneous)."""
    n = depth.numel() // (h * w)
    pix = pixel_grid(h, w, depth.dtype, depth.device).unsqueeze(0)            # [1,3,hw]
    rays = torch.matmul(inv_K[:, :3, :3], pix)                                 # [N|1,3,hw]
    pts = depth.reshape(n, 1, -1) * rays
    ones = torch.ones(n, 1, h * w, dtype=depth.dtype, device=depth.device)
    return torch.cat([pts, ones], 1)


def project(points, K, T, h, w, eps=1e-7):
    """movedepth/layers.py:601-621 -- camera points [N,4,hw] -> grid_sample coords [N,h,w,2]
    (normalised for align_corners=True).  No z>0 test, eps added to z."""
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    uv = cam[:, :2, :] / (cam[:, 2:3, :] + eps)
    uv = uv.reshape(-1, 2, h, w).permute(0, 2, 3, 1).clone()
    uv[..., 0] /= w - 1
    uv[..., 1] /= h - 1
    return (uv - 0.5) * 2


# ----------------------------------------------------------------------------
# depth hypotheses around the mono prior
# ----------------------------------------------------------------------------
def depth_hypotheses(prior_depth, ndepth, scale_fac, z_trans=None, kind="inverse"):
    """movedepth/layers.py:256-284 (`schedule_depth_rangev2`, z_trans=None) and 370-398
    (`schedule_depth_range_zv2`, z_trans [B,1,1,1]).  prior [B,1,h,w] -> [B,D,h,w].
    Index 0 is the FAR end for positive scale (d_max), index D-1 the near end."""
    with torch.no_grad():
        s = scale_fac if z_trans is None else scale_fac * z_trans
        d_lo = prior_depth / (1 + s)
        d_hi = prior_depth * (1 + s)
        _, _, h, w = prior_depth.shape
        itv = torch.arange(0, ndepth, dtype=prior_depth.dtype, device=prior_depth.device)
        itv = itv.reshape(1, -1, 1, 1).repeat(1, 1, h, w) / (ndepth - 1)
        if kind == "inverse":
            return 1 / (1 / d_hi + (1 / d_lo - 1 / d_hi) * itv)
        if kind == "linear":
            return d_lo + (d_hi - d_lo) * itv
        if kind == "log":       # layers.py:276-282 / 389-395: affine map of a fixed 0.1 -> 1 geometric ramp (fp32 loop)
            ramp = torch.stack([torch.exp(torch.log(torch.tensor([0.1])) + torch.log(torch.tensor([1 / 0.1])) *
                                          torch.tensor([float(k)]) / (ndepth - 1)) for k in range(ndepth)]).reshape(1, -1, 1, 1)
            return d_lo + (d_hi - d_lo) * ramp.to(prior_depth.dtype)
        raise NotImplementedError(kind)


def hypothesis_ratios(ndepth, s, dtype=torch.float64):
    """Separable form: hypotheses = prior * ratio[b, k],
    ratio = 1 / (1/(1+s) + ((1+s) - 1/(1+s)) * k/(D-1)).  `s` is [B] (or scalar)."""
    s = torch.as_tensor(s, dtype=dtype).reshape(-1, 1)
    k = torch.arange(ndepth, dtype=dtype).reshape(1, -1) / (ndepth - 1)
    return 1 / (1 / (1 + s) + ((1 + s) - 1 / (1 + s)) * k)


# ----------------------------------------------------------------------------
# cost volume
# ----------------------------------------------------------------------------
def cost_volume(ref, src, K, invK, hyps, pose):
    """movedepth/layers.py:778-794 (`generate_costvol`).

    ref, src [B,C,h,w]; K, invK [B,4,4] (the trainer passes scale-2 intrinsics,
    trainer.py:353); hyps [B,D,h,w]; pose [B,1,4,4] (ref->src).  Returns the
    reference-layout volume [B,D,C,h,w] = bilinear_zeros(src, uv) * ref.  The
    sampling grid is built under no_grad; gradients reach ref and src only.
    """
    B, C, h, w = ref.shape
    D = hyps.shape[1]
    out = []
    for b in range(B):
        with torch.no_grad():
            pts = backproject(hyps[b:b + 1], invK[b:b + 1], h, w)             # [D,4,hw]
            grid = project(pts, K[b:b + 1], pose[b:b + 1, 0], h, w)           # [D,h,w,2]
        warped = F.grid_sample(src[b:b + 1].expand(D, C, h, w), grid, mode="bilinear",
                               padding_mode="zeros", align_corners=True)
        out.append(warped * ref[b:b + 1])
    return torch.stack(out, 0)


def group_correlation(cost, groups):
    """movedepth/trainer.py:359 -- [B,D,C,h,w] -> [B,D,G,h,w]; group g averages channels
    {g, g+G, g+2G, ...} (reshape(B,D,C/G,G,h,w).mean(2))."""
    B, D, C, h, w = cost.shape
    return cost.reshape(B, D, C // groups, groups, h, w).mean(2)


def fuse_views(grouped_list, eval_axis=False):
    """movedepth/trainer.py:349-363 (and evaluate_depth.py:234-242 when eval_axis).

    Per view: weight = max_G softmax_G(mean_D volume) (trainer) -- evaluate_depth uses
    mean over G / softmax over D instead (`cost_vols.mean(2)`); accumulated volume is
    sum(w*vol)/(1e-8+sum w)."""
    wsum = 1e-8
    acc = 0
    for vol in grouped_list:
        wgt = torch.softmax(vol.mean(2 if eval_axis else 1), dim=1).max(1)[0]   # [B,h,w]
        wsum = wsum + wgt
        acc = acc + wgt[:, None, None] * vol
    return acc / wsum[:, None, None]


# ----------------------------------------------------------------------------
# depth regression from the regularised volume
# ----------------------------------------------------------------------------
def entropy(volume, dim, keepdim=False):
    """movedepth/layers.py:862-863."""
    return torch.sum(-volume * volume.clamp(1e-9, 1.).log(), dim=dim, keepdim=keepdim)


def localmax(prob, radius, nbins, inv_a, inv_b):
    """movedepth/layers.py:796-812.  prob [B,D,h,w]; inv_a = 1/hyps[:, -1], inv_b = 1/hyps[:, 0].

    argmax over D, window of clamped indices i*-r..i*+r (clamped duplicates are counted
    twice), soft index = sum(idx*p)/(1e-6+sum p), depth = 1/(inv_a + idx/(D-1)*(inv_b-inv_a)).
    NOTE the orientation: soft index i maps to hypothesis D-1-i (reference behaviour).
    """
    top = torch.argmax(prob, 1, keepdim=True)
    num = 0
    den = 1e-6
    for k in range(-radius, radius + 1):
        idx = (top + k).clamp(0, nbins - 1)
        p = torch.gather(prob, 1, idx)
        num = num + idx * p
        den = den + p
    soft = (num / den) / (nbins - 1)
    return 1 / (inv_a + soft[:, 0] * (inv_b - inv_a))


def convex_upsample(depth, mask, scale=2):
    """movedepth/layers.py:200-214.  depth [B,h,w] or [B,1,h,w]; mask [B,9*f*f,h,w], f=2**scale.
    softmax over the 9 taps, zero-padded 3x3 unfold of depth, pixel-shuffle to [B,f*h,f*w]."""
    if depth.dim() == 3:
        depth = depth.unsqueeze(1)
    B, _, h, w = depth.shape
    f = 2 ** scale
    m = torch.softmax(mask.view(B, 9, f, f, h, w), dim=1)
    nb = F.unfold(depth, [3, 3], padding=1).view(B, 9, 1, 1, h, w)
    up = (m * nb).sum(1)                               # [B,f,f,h,w]
    return up.permute(0, 3, 1, 4, 2).reshape(B, f * h, f * w)


# ----------------------------------------------------------------------------
# photometric loss pieces
# ----------------------------------------------------------------------------
def warp_image(img, depth, K, invK, T):
    """movedepth/trainer.py:501-507 / 519-529 / 575-580: backproject -> project ->
    grid_sample(border, bilinear, align_corners=True).  img [B,3,H,W], depth [B,1,H,W] or [B,H,W]."""
    B, _, H, W = img.shape
    pts = backproject(depth, invK, H, W)
    grid = project(pts, K, T, H, W)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=True), grid


def ssim(x, y):
    """movedepth/layers.py:646-677 -- 3x3 mean-filter SSIM *loss* map, reflection padded,
    clamp((1-SSIM)/2, 0, 1), per channel."""
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sig_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sig_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sig_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + c1) * (2 * sig_xy + c2)
    d = (mu_x ** 2 + mu_y ** 2 + c1) * (sig_x + sig_y + c2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def reprojection_loss(pred, target, ssim_lw=0.85, no_ssim=False):
    """movedepth/trainer.py:535-550 -- [B,3,H,W] x2 -> [B,1,H,W]."""
    l1 = (target - pred).abs().mean(1, True)
    if no_ssim:
        return l1
    return ssim_lw * ssim(pred, target).mean(1, True) + (1 - ssim_lw) * l1


def smooth_loss(disp, img):
    """movedepth/layers.py:630-643 -- edge-aware first-order smoothness (two means added)."""
    gdx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    gdy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs()
    gix = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True)
    giy = (img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True)
    return (gdx * torch.exp(-gix)).mean() + (gdy * torch.exp(-giy)).mean()


def box_mask(img, box_hw, xy=None):
    """movedepth/layers.py:52-69 (`random_image_mask`): zero a box of size box_hw in img.
    Returns (masked image, mask) where mask is 1 OUTSIDE the box and 0 inside.  When `xy`
    is None the position is drawn exactly as the reference does (np.random.randint, x first)."""
    fh, fw = box_hw
    _, _, h, w = img.shape
    if fh == h and fw == w:
        return img, None
    if xy is None:
        x = np.random.randint(0, w - fw)
        y = np.random.randint(0, h - fh)
    else:
        x, y = xy
    m = torch.ones_like(img)
    m[:, :, y:y + fh, x:x + fw] = 0.0
    return img * m, m


def upsampled_depth(disp, height, width, min_depth, max_depth):
    """movedepth/trainer.py:512-515: bilinear (align_corners=False) upsampling of a sigmoid disparity to full
    resolution, then disp_to_depth.  [B,1,hs,ws] -> [B,1,H,W]."""
    up = F.interpolate(disp, [height, width], mode="bilinear", align_corners=False)
    return disp_to_depth(up, min_depth, max_depth)[1]


def normalized_smooth_loss(disp, img):
    """movedepth/trainer.py:712-714: mean-normalised disparity, then the edge-aware smoothness."""
    mean = disp.mean(2, True).mean(3, True)
    return smooth_loss(disp / (mean + 1e-7), img)


def masked_consistency(depth_aug, depth, aug_mask, weight):
    """movedepth/trainer.py:398-400: smooth-L1 (mean) between the two low-resolution depth maps over the pixels where the
    bilinearly resized (align_corners=True) box mask is non-zero; `weight` = mask_lw * mask_lw (applied twice there)."""
    sel = F.interpolate(aug_mask, list(depth_aug.shape[1:]), mode="bilinear", align_corners=True).sum(1).to(torch.bool)
    return F.smooth_l1_loss(depth_aug[sel], depth[sel], reduction="mean") * weight
