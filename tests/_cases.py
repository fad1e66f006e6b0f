"""Seeded test inputs shared by the golden generator, the oracle tests and the GPU parity tests.

Everything here is a pure function of fixed seeds (torch CPU generators), so the same inputs
are rebuilt in the build container (where the reference produced tests/golden/*.npz) and on
the GPU box.  No reference code is imported here.
"""
import math

import torch
import torch.nn.functional as F


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def intrinsics(batch, h, w):
    """KITTI normalised intrinsics scaled to an h x w image; returns (K, inv_K) as [B,4,4] fp32."""
    K = torch.tensor([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float64)
    K[0] *= w
    K[1] *= h
    return K.float().repeat(batch, 1, 1), torch.linalg.pinv(K).float().repeat(batch, 1, 1)


def rigid(axisangle, translation):
    """fp64 Rodrigues -> fp32 [4,4] (test input only)."""
    a = torch.tensor(axisangle, dtype=torch.float64)
    th = a.norm()
    Kx = torch.tensor([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=torch.float64)
    R = torch.eye(3, dtype=torch.float64)
    if th > 0:
        R = R + math.sin(th) / th * Kx + (1 - math.cos(th)) / th ** 2 * (Kx @ Kx)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = R
    T[:3, 3] = torch.tensor(translation, dtype=torch.float64)
    return T.float()


def smooth_noise(shape, seed, factor=2, normal=True):
    b, c, h, w = shape
    g = _gen(seed)
    lo = (torch.randn if normal else torch.rand)(b, c, max(2, h // factor), max(2, w // factor), generator=g)
    return F.interpolate(lo, size=(h, w), mode="bicubic", align_corners=True).contiguous()


def ground_prior(batch, h, w, seed):
    """Ground-plane-like depth profile, network-scale units."""
    y = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1).expand(batch, 1, h, w)
    base = (1.2 / ((y - 0.45 * h).clamp(min=1.0) / h * 1.92)).clamp(0.5, 60.0)
    return (base * (1 + 0.1 * torch.rand(batch, 1, h, w, generator=_gen(seed)))).contiguous()


def ratios(D, s):
    """hypothesis = prior * ratio[b,k], fp64 -> fp32 [B,D]."""
    s = torch.as_tensor(s, dtype=torch.float64).reshape(-1, 1)
    k = torch.arange(D, dtype=torch.float64).reshape(1, -1) / (D - 1)
    return (1 / (1 / (1 + s) + ((1 + s) - 1 / (1 + s)) * k)).float()


# ---------------------------------------------------------------- operator cases
def case_hypotheses():
    g = _gen(11)
    return dict(prior=0.5 + 20 * torch.rand(2, 1, 8, 16, generator=g), D=8, fac=0.3,
                z_trans=torch.tensor([30 * -0.035, 30 * 0.02]).view(2, 1, 1, 1))


POSES = {
    "identity": ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0)),
    "forward": ((0.002, 0.010, 0.001), (0.002, 0.001, -0.035)),
    "sideways": ((0.001, -0.004, 0.002), (0.10, 0.0, -0.01)),
    "stress": ((0.006, 0.030, 0.003), (0.02, 0.005, -0.10)),
}
COSTVOL_CASES = list(POSES)


def case_costvol(name, B=2, C=32, h=8, w=40, D=4, seed=20):
    aa, tr = POSES[name]
    K, invK = intrinsics(B, h, w)
    prior = ground_prior(B, h, w, seed + 1)
    two = [rigid(aa, tr), rigid([-x for x in aa], [-0.5 * x for x in tr])]
    pose = torch.stack([two[i % 2] for i in range(B)])
    # velocity-guided range (zv2): s_b = depth_bin_fac * z_scale * T[b,2,3]; identity uses the fixed v2 range
    s = [0.3 if name == "identity" else 0.3 * 30 * float(pose[i, 2, 3]) for i in range(B)]
    if name == "forward":
        s[-1] = 0.3           # mix the v2 (fixed 0.3) and zv2 schedules in one batch
    rat = ratios(D, s)
    hyps = (prior * rat.view(B, D, 1, 1)).contiguous()
    return dict(B=B, C=C, h=h, w=w, D=D, K=K, invK=invK, prior=prior, ratio=rat, hyps=hyps,
                pose=pose.unsqueeze(1).contiguous(),
                ref=smooth_noise((B, C, h, w), seed + 2), src=smooth_noise((B, C, h, w), seed + 3),
                gvol=torch.randn(B, D, C, h, w, generator=_gen(seed + 4)))


def case_localmax(B=2, D=8, h=8, w=16):
    g = _gen(31)
    prob = torch.softmax(2.0 * torch.randn(B, D, h, w, generator=g), 1)
    inv_a = 0.02 + torch.rand(B, h, w, generator=g)
    inv_b = inv_a * (1.2 + torch.rand(B, h, w, generator=g))
    idx = (torch.arange(B * h * w) % D).view(B, 1, h, w)
    onehot = torch.zeros(B, D, h, w).scatter_(1, idx, 1.0)
    return dict(prob=prob, inv_a=inv_a, inv_b=inv_b, onehot=onehot, D=D)


def case_convex(B=2, h=8, w=16):
    g = _gen(41)
    return dict(depth=0.5 + 10 * torch.rand(B, h, w, generator=g), mask=torch.randn(B, 144, h, w, generator=g))


def case_images(B=2, H=16, W=24):
    g = _gen(51)
    x = torch.rand(B, 3, H, W, generator=g)
    y = (x + 0.2 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    return dict(x=x, y=y, disp=torch.rand(B, 1, H, W, generator=g))


def case_pose():
    g = _gen(61)
    aa = 0.05 * torch.randn(4, 1, 3, generator=g)
    aa[1] = 0.0                                 # exercises the angle + 1e-7 guard
    return dict(aa=aa, tr=0.1 * torch.randn(4, 1, 3, generator=g))


def case_warp(B=2, H=16, W=24):
    g = _gen(71)
    K, invK = intrinsics(B, H, W)
    T = torch.stack([rigid(*POSES["forward"]), rigid((0.01, -0.02, 0.005), (0.3, -0.05, 0.2))])
    return dict(B=B, H=H, W=W, K=K, invK=invK, T=T, img=torch.rand(B, 3, H, W, generator=g),
                depth=1.0 + 5 * torch.rand(B, 1, H, W, generator=g))


def case_disp_pyramid(B=2, H=32, W=64):
    """Sigmoid-like disparities at the 4 decoder scales + gradient seeds (disp -> full-res depth kernel)."""
    g = _gen(91)
    return dict(B=B, H=H, W=W, disp=[torch.rand(B, 1, H // 2 ** s, W // 2 ** s, generator=g) for s in range(4)],
                gdepth=torch.randn(B, 1, H, W, generator=g), img=torch.rand(B, 3, H, W, generator=g))


def case_masked(B=2, h=12, w=20, H=48, W=80):
    """Two low-resolution depth maps and the augmentation boxes (x, y) of the masked-consistency term."""
    g = _gen(101)
    a = 1.0 + 8 * torch.rand(B, h, w, generator=g)
    return dict(B=B, h=h, w=w, H=H, W=W, a=a, b=a + 1.5 * torch.randn(B, h, w, generator=g),
                boxes=[(0, 0), (17, 9), (W - W // 3 - 1, H - H // 3 - 1), (31, 2)])


def case_frames(H=32, W=96, Hn=75, Wn=250, seed=9):
    """Two synthetic 'decoded' uint8 frames at a native resolution (data pipeline tests), numpy generator."""
    import numpy as np
    rng = np.random.default_rng(111)
    frames = {}
    for f in (0, -1):
        walk = np.cumsum(rng.normal(0, 6, (Hn, Wn, 3)), 1) + np.cumsum(rng.normal(0, 4, (Hn, Wn, 3)), 0)
        frames[f] = np.clip(walk + 128, 0, 255).astype(np.uint8)
    return dict(H=H, W=W, Hn=Hn, Wn=Wn, seed=seed, frames=frames)


# ---------------------------------------------------------------- whole-step cases
STEP_CASES = {
    "r18_2f": dict(H=64, W=96, D=8, B=2, frame_ids=[0, -1], epoch=0, arch=18),
    "r18_2f_z": dict(H=64, W=96, D=8, B=2, frame_ids=[0, -1], epoch=9, arch=18),
    "r18_3f": dict(H=64, W=96, D=16, B=2, frame_ids=[0, -1, 1], epoch=0, arch=18),
    "r50_3f": dict(H=64, W=96, D=8, B=1, frame_ids=[0, -1, 1], epoch=9, arch=50),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case): R18 2-frame 192x640 D=16 batch 2, velocity-guided range
    "c1": dict(H=192, W=640, D=16, B=2, frame_ids=[0, -1], epoch=9, arch=18, slim=True),
}
EVAL_CASE = dict(H=64, W=96, D=8, B=2, frame_ids=[0, -1], epoch=9, arch=18)      # inference path (evaluate_depth.py:181-253)
GRAD_PROBES = [("pose", "net.3.bias"), ("reg3d", "prob.weight"), ("mask_cnn", "head_convs.weight"),
               ("mono_depth", "decoder.10.conv.bias"), ("up", "upsample_mask.2.weight"),
               ("mvs_encoder", "out.weight")]


def step_options(cfg):
    from oracle.step import default_options
    return default_options(height=cfg["H"], width=cfg["W"], num_depth_bins=cfg["D"], batch_size=cfg["B"],
                           frame_ids=list(cfg["frame_ids"]), matching_ids=[0, -1], res_arch=cfg.get("arch", 18))


def step_inputs(cfg):
    """(inputs dict, noise list, (x, y) of the augmentation box)."""
    from oracle.step import synthetic_inputs
    opt = step_options(cfg)
    inputs = synthetic_inputs(opt, cfg["B"], seed=1, smooth=True)
    g = _gen(81)
    noise = [torch.randn(cfg["B"], 1, cfg["H"], cfg["W"], generator=g) for _ in range(4)]
    fh, fw = cfg["H"] // 3, cfg["W"] // 3
    mask_xy = (int(torch.randint(0, cfg["W"] - fw, (1,), generator=g)), int(torch.randint(0, cfg["H"] - fh, (1,), generator=g)))
    return inputs, noise, mask_xy
