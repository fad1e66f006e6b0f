"""Host-side mirror of the reference's geometry / cost-volume / loss operator library.

Same names, argument meaning and return shapes as `movedepth/layers.py` of the reference
(imported by name at movedepth/trainer.py:22-24 and evaluate_depth.py:15-16), so caller code
drops in unchanged.  The dense operators dispatch to the hand-written sm_100a kernels in
libmovedepth_b200.so (movedepth_b200.ops); small glue (pose algebra, hypothesis tables) is
plain tensor code on the caller's device.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from . import precision as PR


# ------------------------------------------------------------------ disparity / pose algebra
def disp_to_depth(disp, min_depth, max_depth):
    """Sigmoid disparity -> (scaled disparity, depth).  Reference: movedepth/layers.py:400-409."""
    inv_far, inv_near = 1.0 / max_depth, 1.0 / min_depth
    scaled = inv_far + (inv_near - inv_far) * disp
    return scaled, 1.0 / scaled


def rot_from_axisangle(vec):
    """Axis-angle [B,1,3] -> homogeneous rotation [B,4,4] (Rodrigues, angle+1e-7 guard).
    Reference: movedepth/layers.py:479-518."""
    theta = torch.norm(vec, 2, 2, True)
    unit = vec / (theta + 1e-7)
    ca, sa = torch.cos(theta), torch.sin(theta)
    t = 1 - ca
    ux, uy, uz = unit[..., 0:1], unit[..., 1:2], unit[..., 2:3]
    tx, ty, tz = ux * t, uy * t, uz * t
    rows = [
        ux * tx + ca, ux * ty - uz * sa, uz * tx + uy * sa,
        ux * ty + uz * sa, uy * ty + ca, uy * tz - ux * sa,
        uz * tx - uy * sa, uy * tz + ux * sa, uz * tz + ca,
    ]
    R3 = torch.cat(rows, 2).reshape(-1, 3, 3)
    R = torch.zeros(vec.shape[0], 4, 4, dtype=vec.dtype, device=vec.device)
    R[:, :3, :3] = R3
    R[:, 3, 3] = 1
    return R


def get_translation_matrix(translation_vector):
    """[B,1,3] -> [B,4,4].  Reference: movedepth/layers.py:464-477."""
    n = translation_vector.shape[0]
    T = torch.eye(4, dtype=translation_vector.dtype, device=translation_vector.device).repeat(n, 1, 1)
    T[:, :3, 3] = translation_vector.reshape(n, 3)
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    """Pose-net outputs -> 4x4 camera transform.  Reference: movedepth/layers.py:412-429.  CUDA fp32 inputs take the fused
    kernel (one launch per direction instead of ~50 elementwise ones)."""
    if axisangle.is_cuda and axisangle.dtype == torch.float32 and translation.dtype == torch.float32 and axisangle.numel() == 3 * axisangle.shape[0]:
        from . import ops
        return ops.pose_matrix(axisangle, translation, invert)
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = -t
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


# ------------------------------------------------------------------ depth hypotheses
def hypothesis_ratios(ndepth, scale, device, type="inverse"):
    """Separable form of the hypothesis schedule: hypotheses = prior * ratio[b, k].

    `scale` is a python float or a tensor with one value per batch item (depth_bin_fac, or
    depth_bin_fac * z_scale * T[b,2,3] for the velocity-guided schedule).  Returns [B,D] fp32.
    Derived from movedepth/layers.py:261-267 / 375-381: d_min = c/(1+s), d_max = c(1+s),
    1/d_k = 1/d_max + (1/d_min - 1/d_max) k/(D-1)."""
    s = torch.as_tensor(scale, dtype=torch.float32, device=device).reshape(-1, 1)
    k = torch.arange(ndepth, dtype=torch.float32, device=device).reshape(1, -1) / (ndepth - 1)
    if type == "inverse":
        return 1.0 / (1.0 / (1 + s) + ((1 + s) - 1.0 / (1 + s)) * k)
    if type == "linear":
        return 1.0 / (1 + s) + ((1 + s) - 1.0 / (1 + s)) * k
    if type == "log":       # affine map of a fixed 0.1 -> 1 geometric ramp (layers.py:276-282)
        return 1.0 / (1 + s) + ((1 + s) - 1.0 / (1 + s)) * _log_ramp(ndepth, device).reshape(1, -1)
    raise NotImplementedError(type)


def _log_ramp(ndepth, device):
    """itv_k = exp(log(0.1) + log(10) * k / (D-1)), evaluated in fp32 like the reference's FloatTensor loop."""
    k = torch.arange(ndepth, dtype=torch.float32)
    lo, span = torch.log(torch.tensor([0.1])), torch.log(torch.tensor([1 / 0.1]))
    return torch.exp(lo + span * k / (ndepth - 1)).to(device)


def _schedule(prior_depth, ndepth, s, type):
    with torch.no_grad():
        lo = prior_depth / (1 + s)
        hi = prior_depth * (1 + s)
        k = torch.arange(ndepth, dtype=prior_depth.dtype, device=prior_depth.device).reshape(1, -1, 1, 1) / (ndepth - 1)
        if type == "inverse":
            return 1.0 / (1.0 / hi + (1.0 / lo - 1.0 / hi) * k)
        if type == "linear":
            return lo + (hi - lo) * k
        if type == "log":
            return lo + (hi - lo) * _log_ramp(ndepth, prior_depth.device).reshape(1, -1, 1, 1)
        raise NotImplementedError(type)


def schedule_depth_rangev2(prior_depth, ndepth, scale_fac, type="inverse"):
    """[B,1,h,w] prior -> [B,D,h,w] hypotheses in [c/(1+s), c(1+s)], index 0 = far end.
    Reference: movedepth/layers.py:256-284."""
    return _schedule(prior_depth, ndepth, scale_fac, type)


def schedule_depth_range_zv2(prior_depth, ndepth, scale_fac, z_trans, type="inverse"):
    """Velocity-guided range: s = scale_fac * z_trans (per batch item, [B,1,1,1]).
    Reference: movedepth/layers.py:370-398."""
    return _schedule(prior_depth, ndepth, scale_fac * z_trans, type)


# ------------------------------------------------------------------ pinhole geometry modules
class BackprojectDepth(nn.Module):
    """depth [B,1,H,W] (+ inv_K [B|1,4,4]) -> homogeneous camera points [B,4,HW].
    Reference: movedepth/layers.py:556-586.  The fused kernels do this per pixel in registers;
    the module exists for callers that want the point cloud itself."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32),
                                torch.arange(width, dtype=torch.float32), indexing="ij")
        pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width)], 0)
        # registered as frozen parameters like the reference (they show up in .parameters())
        self.pix_coords = nn.Parameter(pix.unsqueeze(0).repeat(batch_size, 1, 1), requires_grad=False)
        self.ones = nn.Parameter(torch.ones(batch_size, 1, height * width), requires_grad=False)

    def forward(self, depth, inv_K):
        rays = torch.matmul(inv_K[:, :3, :3], self.pix_coords)
        pts = depth.reshape(self.batch_size, 1, -1) * rays
        return torch.cat([pts, self.ones], 1)


class Project3D(nn.Module):
    """points [B,4,HW] -> grid_sample coordinates [B,H,W,2] (align_corners=True convention).
    Reference: movedepth/layers.py:589-621."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T):
        P = torch.matmul(K, T)[:, :3, :]
        cam = torch.matmul(P, points)
        uv = cam[:, :2] / (cam[:, 2:3] + self.eps)
        uv = uv.reshape(self.batch_size, 2, self.height, self.width).permute(0, 2, 3, 1)
        scale = uv.new_tensor([self.width - 1, self.height - 1])
        return (uv / scale - 0.5) * 2


# ------------------------------------------------------------------ cost volume
def generate_costvol(ref, src, K, invK, depth_priors, pose, num_depth_bins, backprojector=None, projector=None):
    """Reference-layout homography-warp volume: [B,D,C,h,w] = bilinear_zeros(src, uv(d)) * ref.
    Same signature as movedepth/layers.py:778-794 (`backprojector`/`projector` are accepted and
    ignored: the kernel does the projection per pixel).  The trainer's fast path uses
    `fused_group_costvol` instead and never materialises this tensor."""
    assert depth_priors.shape[1] == num_depth_bins
    return ops.costvol_full(ref, src, depth_priors, K, invK, pose[:, 0])


def fused_group_costvol(ref, src, K, invK, pose, prior, ratio, groups=16, layout=ops.LAYOUT_BGDHW, flags=0):
    """generate_costvol + group mean (movedepth/trainer.py:358-359) in one kernel.
    prior [B,1,h,w], ratio [B,D] -> grouped volume, logical shape [B,G,D,h,w]."""
    return ops.costvol_grouped(ref, src, K, invK, pose, prior=prior, ratio=ratio, groups=groups, layout=layout,
                               flags=flags)


# ------------------------------------------------------------------ depth regression
def entropy(volume, dim, keepdim=False):
    """Reference: movedepth/layers.py:862-863."""
    return torch.sum(-volume * volume.clamp(1e-9, 1.).log(), dim=dim, keepdim=keepdim)


def localmax(cost_prob, radius, casbin, min_depth_inverse, max_depth_inverse):
    """Local soft-argmax depth from a probability volume [B,D,h,w] -> [B,h,w].
    Reference: movedepth/layers.py:796-812.  The kernel takes logits; log(prob) has the same
    softmax, so this entry point feeds it log-probabilities."""
    assert cost_prob.shape[1] == casbin
    _, _, depth = ops.regress_depth(torch.log(cost_prob.clamp_min(1e-38)), min_depth_inverse, max_depth_inverse, radius)
    return depth


def convex_upsample(depth, mask, scale=2):
    """Reference: movedepth/layers.py:200-214."""
    return ops.convex_upsample(depth, mask, scale)


class convex_upsample_layer(nn.Module):
    """Mask head (3x3 conv -> ReLU -> 1x1 conv, no biases) + convex upsampling by 2**scale.
    Reference: movedepth/layers.py:184-198."""

    def __init__(self, feature_dim, scale=2):
        super().__init__()
        self.scale = scale
        self.upsample_mask = nn.Sequential(
            PR.Conv2d(feature_dim, 64, 3, stride=1, padding=1, bias=False),
            nn.ReLU(inplace=True),
            PR.Conv2d(64, (2 ** scale) ** 2 * 9, 1, stride=1, padding=0, bias=False))

    def forward(self, depth, feat):
        return convex_upsample(depth, self.upsample_mask(feat), self.scale)


# ------------------------------------------------------------------ decoder building blocks
class Conv3x3(nn.Module):
    """Reflection (or zero) pad 1 + 3x3 conv.  Reference: movedepth/layers.py:537-553."""

    def __init__(self, in_channels, out_channels, use_refl=True):
        super().__init__()
        self.pad = nn.ReflectionPad2d(1) if use_refl else nn.ZeroPad2d(1)
        self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3)

    def forward(self, x):
        return self.conv(self.pad(x))


class ConvBlock(nn.Module):
    """Conv3x3 + ELU.  Reference: movedepth/layers.py:521-534."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU(inplace=True)

    def forward(self, x):
        return self.nonlin(self.conv(x))


def upsample(x):
    """Nearest x2.  Reference: movedepth/layers.py:624-627."""
    return F.interpolate(x, scale_factor=2, mode="nearest")


# ------------------------------------------------------------------ photometric pieces
def get_smooth_loss(disp, img):
    """Edge-aware first-order smoothness.  Reference: movedepth/layers.py:630-643.  Device tensors go through the
    `mvd_smooth_loss` kernels; the tensor formula below serves host-side callers."""
    if disp.is_cuda and disp.dim() == 4 and disp.shape[1] == 1 and img.shape[1] == 3 and not img.requires_grad:
        return ops.smooth_loss(disp, img, normalize=False)
    dx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    dy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs()
    ix = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True)
    iy = (img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True)
    return (dx * torch.exp(-ix)).mean() + (dy * torch.exp(-iy)).mean()


class SSIM(nn.Module):
    """3x3 mean-filter SSIM loss map, reflection padded: clamp((1-SSIM)/2, 0, 1).
    Reference: movedepth/layers.py:646-677."""
    C1 = 0.01 ** 2
    C2 = 0.03 ** 2

    def forward(self, x, y):
        xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
        yp = F.pad(y, (1, 1, 1, 1), mode="reflect")
        mx, my = F.avg_pool2d(xp, 3, 1), F.avg_pool2d(yp, 3, 1)
        sx = F.avg_pool2d(xp * xp, 3, 1) - mx * mx
        sy = F.avg_pool2d(yp * yp, 3, 1) - my * my
        sxy = F.avg_pool2d(xp * yp, 3, 1) - mx * my
        num = (2 * mx * my + self.C1) * (2 * sxy + self.C2)
        den = (mx * mx + my * my + self.C1) * (sx + sy + self.C2)
        return torch.clamp((1 - num / den) / 2, 0, 1)


def random_image_mask(img, filter_size):
    """Zero a random filter_size=(fh,fw) box.  Returns (masked image, mask with 0 inside the box).
    Reference: movedepth/layers.py:52-69 (np.random.randint, x drawn before y)."""
    fh, fw = filter_size
    _, _, h, w = img.shape
    if fh == h and fw == w:
        return img, None
    x = np.random.randint(0, w - fw)
    y = np.random.randint(0, h - fh)
    mask = torch.ones_like(img)
    mask[:, :, y:y + fh, x:x + fw] = 0.0
    return img * mask, mask


def compute_depth_errors(gt, pred):
    """Standard depth metrics (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3).
    Reference: movedepth/layers.py:718-736."""
    ratio = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = [(ratio < 1.25 ** k).float().mean() for k in (1, 2, 3)]
    diff = gt - pred
    rmse = torch.sqrt((diff ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = (diff.abs() / gt).mean()
    sq_rel = (diff ** 2 / gt).mean()
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
