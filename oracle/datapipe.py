"""Oracle restatement of the reference's image pre-processing (CPU, numpy integer arithmetic).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows movedepth/datasets/mono_dataset.py:
  * 104-126 `preprocess`: the 4-scale pyramid, every scale resized from the previous one with
    `transforms.Resize(..., interpolation=Image.ANTIALIAS)` (= PIL's LANCZOS resampler), `to_tensor`, colour augmentation;
  * 70-80, 220-223: `transforms.ColorJitter(brightness, contrast, saturation, hue)` on the PIL image;
  * 164, 206: horizontal flip; 209-218: intrinsics per scale.
The arithmetic lives in third-party dependencies that are not vendored: Pillow's `ImagingResample` (Resample.c: double
coefficients normalised to 22-bit fixed point, int32 accumulation with a rounding bias, horizontal pass then vertical pass
through an 8-bit intermediate), `ImageEnhance` (`Image.blend`: float arithmetic truncated to uint8), `Convert.c`
(`rgb2hsv` / `hsv2rgb`, L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16) and torchvision's `adjust_hue`.  The restatements
below are pinned bit-exactly against the Pillow / torchvision of this container by tests/test_datapipe_oracle.py
(random images, and all 2^24 RGB values for the colour conversions).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    return _sinc(x) * _sinc(x / 3.0) if -3.0 <= x < 3.0 else 0.0


def lanczos_coefficients(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = (0, in_size)).
    Returns (bounds int32 [out,2] = (xmin, count), coeffs int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img, out_size, axis):
    """One pass of ImagingResample{Horizontal,Vertical}_8bpc along `axis` of a uint8 array [..., H, W, C]."""
    img = np.moveaxis(img, axis, -1).astype(np.int64)                       # [..., n_in]
    bounds, kk = lanczos_coefficients(img.shape[-1], out_size)
    out = np.empty(img.shape[:-1] + (out_size,), dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (img[..., xmin:xmin + cnt] * kk[xx, :cnt].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def resize_lanczos(img, height, width):
    """uint8 [..., H, W, C] -> [..., height, width, C]: horizontal pass, then vertical pass (Resample.c:ImagingResample)."""
    h, w = img.shape[-3], img.shape[-2]
    if w != width:
        img = _resample_axis(img, width, -2)
    if h != height:
        img = _resample_axis(img, height, -3)
    return img


def pyramid(img, height, width, num_scales=4):
    """mono_dataset.py:104-126: scale i is resized from scale i-1 (scale 0 from the native image)."""
    out = []
    for i in range(num_scales):
        img = resize_lanczos(img, height // 2 ** i, width // 2 ** i)
        out.append(img)
    return out


def to_tensor(img):
    """transforms.ToTensor: uint8 HWC -> float32 CHW / 255."""
    return np.moveaxis(img, -1, -3).astype(np.float32) / np.float32(255.0)


# ----------------------------------------------------------------------------------------------- colour jitter
def _blend(degenerate, img, factor):
    """PIL Image.blend(degenerate, img, factor) (Blend.c): float arithmetic, truncation; clipping only when extrapolating."""
    a = np.float32(factor)
    d, x = degenerate.astype(np.float32), img.astype(np.float32)
    t = (d + a * (x - d)).astype(np.float32)
    if 0.0 <= factor <= 1.0:
        return t.astype(np.uint8)                       # (UINT8) of a value in [0, 255]: truncation
    return np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32))).astype(np.uint8)


def luminance(img):
    """Convert.c rgb2l: L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).astype(np.uint8)


def adjust_brightness(img, factor):
    return _blend(np.zeros_like(img), img, factor)


def adjust_contrast(img, factor):
    mean = int(luminance(img).astype(np.float64).sum() / (img.shape[-3] * img.shape[-2]) + 0.5)     # ImageStat mean of "L"
    return _blend(np.full_like(img, mean), img, factor)


def adjust_saturation(img, factor):
    return _blend(np.repeat(luminance(img)[..., None], 3, -1), img, factor)


def rgb_to_hsv(img):
    """Convert.c rgb2hsv_row."""
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    maxc, minc = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(np.float32)
    safe = np.where(cr == 0, np.float32(1), cr)
    s = cr / np.where(maxc == 0, 1, maxc).astype(np.float32)
    rc, gc, bc = ((maxc - c).astype(np.float32) / safe for c in (r, g, b))
    d = np.float64                                         # `2.0 + rc - bc` is evaluated in double (double literal), then stored to float
    h = np.where(r == maxc, bc - gc, np.where(g == maxc, (2.0 + rc.astype(d) - bc.astype(d)).astype(np.float32),
                                              (4.0 + gc.astype(d) - rc.astype(d)).astype(np.float32))).astype(np.float32)
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(np.float32)
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    gray = minc == maxc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def hsv_to_rgb(img):
    """Convert.c hsv2rgb."""
    h, s, v = (img[..., i].astype(np.float32) for i in range(3))
    hh = (h.astype(np.float64) * 6.0 / 255.0)
    i = np.floor(hh.astype(np.float32)).astype(np.int32)
    f = (hh - i.astype(np.float32).astype(np.float64)).astype(np.float32)
    fs = (s.astype(np.float64) / 255.0).astype(np.float32)

    def rnd(x):                                           # C round(): half away from zero (arguments are >= 0 here)
        return np.floor(x + 0.5).astype(np.int32)
    p = rnd(v.astype(np.float64) * (1.0 - fs.astype(np.float64)))
    q = rnd(v.astype(np.float64) * (1.0 - fs.astype(np.float64) * f.astype(np.float64)))
    t = rnd(v.astype(np.float64) * (1.0 - fs.astype(np.float64) * (1.0 - f.astype(np.float64))))
    p, q, t = (np.clip(x, 0, 255) for x in (p, q, t))
    vi = v.astype(np.int32)
    sel = i % 6
    r = np.choose(sel, [vi, q, p, p, t, vi])
    g = np.choose(sel, [t, vi, vi, q, p, p])
    b = np.choose(sel, [p, p, t, vi, vi, q])
    gray = img[..., 1] == 0
    return np.stack([np.where(gray, vi, r), np.where(gray, vi, g), np.where(gray, vi, b)], -1).astype(np.uint8)


def adjust_hue(img, factor):
    """torchvision F_pil.adjust_hue: HSV, h += uint8(factor * 255) with wrap-around, back to RGB."""
    hsv = rgb_to_hsv(img)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + int(np.uint8(np.int32(factor * 255)))).astype(np.uint8)
    return hsv_to_rgb(hsv)


JITTER_OPS = (adjust_brightness, adjust_contrast, adjust_saturation, adjust_hue)


def color_jitter(img, order, factors):
    """transforms.ColorJitter.forward with the draws of get_params: `order` = permutation of (0 brightness, 1 contrast,
    2 saturation, 3 hue), `factors` = (b, c, s, h)."""
    for k in order:
        img = JITTER_OPS[int(k)](img, float(factors[int(k)]))
    return img


def scaled_intrinsics(K_norm, height, width, num_scales=4):
    """mono_dataset.py:209-218: K rows 0 / 1 times width / height of the scale; inv_K = pinv(K)."""
    out = []
    for s in range(num_scales):
        K = np.array(K_norm, dtype=np.float32).copy()
        K[0, :] *= width // 2 ** s
        K[1, :] *= height // 2 ** s
        out.append((K, np.linalg.pinv(K)))
    return out
