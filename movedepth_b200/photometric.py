"""Photometric reprojection primitives used by the trainer's loss assembly.

`reprojection_loss`  = BackprojectDepth -> Project3D -> grid_sample(border) -> SSIM/L1 mix for one
source frame (movedepth/trainer.py:519-529 + 535-550) as ONE fused kernel (forward) and one
(backward: d/d depth and d/d pose); `identity_loss` is the same comparison without the warp
(trainer.py:689-693).  Device tensors only.
"""
from . import ops
from .layers import SSIM

_ssim = SSIM()


def reprojection_loss(depth, src, tgt, K, inv_K, T, ssim_w=0.85):
    """-> (loss [B,1,H,W], warped [B,3,H,W])."""
    return ops.photometric_loss(depth, src, tgt, K, inv_K, T, ssim_w)


def identity_loss(src, tgt, ssim_w=0.85):
    return ops.photometric_identity(src, tgt, ssim_w)


def photo_error(pred, target, ssim_w=0.85):
    """Public `compute_reprojection_loss(pred, target)` form: the caller supplies the prediction.
    Differentiable w.r.t. pred (tensor code); the trainer's own losses go through the fused kernel."""
    if not pred.requires_grad:
        return ops.photometric_identity(pred, target, ssim_w)
    l1 = (target - pred).abs().mean(1, True)
    if ssim_w == 0:
        return l1
    return ssim_w * _ssim(pred, target).mean(1, True) + (1 - ssim_w) * l1
