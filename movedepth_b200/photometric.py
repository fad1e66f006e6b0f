"""Photometric reprojection primitives used by the trainer's loss assembly.

`reprojection_loss`  = BackprojectDepth -> Project3D -> grid_sample(border) -> SSIM/L1 mix for one
source frame (movedepth/trainer.py:519-529 + 535-550); `identity_loss` is the same comparison
without the warp (trainer.py:689-693).  Device tensors only.
"""
import torch
import torch.nn.functional as F

from .layers import SSIM

_ssim = SSIM()


def _pixel_rays(inv_K, H, W):
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=inv_K.device),
                            torch.arange(W, dtype=torch.float32, device=inv_K.device), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=inv_K.device)], 0)
    return torch.matmul(inv_K[:, :3, :3], pix.unsqueeze(0))           # [B,3,HW]


def warp(src, depth, K, inv_K, T):
    """Inverse-warp `src` [B,3,H,W] into the target view with per-pixel `depth` ([B,1,H,W] or [B,H,W])."""
    B, _, H, W = src.shape
    rays = _pixel_rays(inv_K, H, W)
    pts = depth.reshape(B, 1, -1) * rays
    pts = torch.cat([pts, torch.ones(B, 1, H * W, device=src.device)], 1)
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, pts)
    uv = cam[:, :2] / (cam[:, 2:3] + 1e-7)
    uv = uv.reshape(B, 2, H, W).permute(0, 2, 3, 1)
    grid = (uv / uv.new_tensor([W - 1, H - 1]) - 0.5) * 2
    return F.grid_sample(src, grid, mode="bilinear", padding_mode="border", align_corners=True)


def photo_error(pred, target, ssim_w=0.85):
    """ssim_w * mean_c SSIM + (1-ssim_w) * mean_c L1 -> [B,1,H,W]; ssim_w == 0 -> L1 only."""
    l1 = (target - pred).abs().mean(1, True)
    if ssim_w == 0:
        return l1
    return ssim_w * _ssim(pred, target).mean(1, True) + (1 - ssim_w) * l1


def reprojection_loss(depth, src, tgt, K, inv_K, T, ssim_w=0.85):
    warped = warp(src, depth, K, inv_K, T)
    return photo_error(warped, tgt, ssim_w), warped


def identity_loss(src, tgt, ssim_w=0.85):
    return photo_error(src, tgt, ssim_w)
