"""Device-side replacement of the reference loader's image pre-processing.

`DevicePipeline` turns decoded uint8 frames that already sit on the GPU into the item dict of
`movedepth/datasets/mono_dataset.py:134-154` -- `("color", f, s)`, `("color_aug", f, s)` for the four scales, `("K", s)`,
`("inv_K", s)` -- doing what `MonoDataset.__getitem__` / `preprocess` (mono_dataset.py:104-126, 156-237) do per item on the
CPU through Pillow and torchvision: optional horizontal flip, the pyramid of LANCZOS (`Image.ANTIALIAS`) resizes, each
scale resized from the previous one, `ToTensor`, and `transforms.ColorJitter` on the uint8 image.  The kernels
(csrc/datapipe.cu) are bit-exact to Pillow / torchvision, so a model sees the same pixels whichever loader feeds it.
Decoding JPEG/PNG files stays outside (nvJPEG or a CPU decoder pool); at > 200 frames/s per GPU the twelve PIL worker
processes of train_movedepth.sh:28 are what limits the reference's input pipeline.
"""
import ctypes
import math
import random

import numpy as np
import torch

from . import _lib, ops

PRECISION_BITS = 32 - 8 - 2
BRIGHTNESS, CONTRAST, SATURATION, HUE = (0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1)     # mono_dataset.py:70-73


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos_taps(in_size, out_size):
    """The coefficient table Pillow builds for one axis (Resample.c: precompute_coeffs + normalize_coeffs_8bpc, LANCZOS,
    support 3): bounds int32 [out,2] = (first tap, count), taps int32 [out,ksize] in 22-bit fixed point."""
    scale = float(in_size) / out_size
    fscale = max(scale, 1.0)
    support = 3.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    taps = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / fscale
    for o in range(out_size):
        center = (o + 0.5) * scale
        first = max(int(center - support + 0.5), 0)
        count = min(int(center + support + 0.5), in_size) - first
        w, total = [], 0.0
        for x in range(count):
            t = (x + first - center + 0.5) * inv
            v = _sinc(t) * _sinc(t / 3.0) if -3.0 <= t < 3.0 else 0.0
            w.append(v)
            total += v
        for x, v in enumerate(w):
            if total != 0.0:
                v = v / total
            taps[o, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[o] = (first, count)
    return bounds, taps


def sample_item_flags(is_train=True, load_pose=False):
    """mono_dataset.py:156-157: (do_color_aug, do_flip) from Python's `random`, colour first."""
    do_color_aug = is_train and random.random() > 0.5
    do_flip = is_train and random.random() > 0.5 and (not load_pose)
    return do_color_aug, do_flip


def sample_color_jitter(generator=None):
    """The draws of torchvision's ColorJitter.get_params (one call = one image): order = randperm(4) over (brightness,
    contrast, saturation, hue), then the four factors, from torch's CPU generator."""
    order = torch.randperm(4, generator=generator)
    f = [float(torch.empty(1).uniform_(lo, hi, generator=generator)) for lo, hi in (BRIGHTNESS, CONTRAST, SATURATION, HUE)]
    return [int(k) for k in order], f


class DevicePipeline:
    def __init__(self, height, width, num_scales=4, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("movedepth_b200 has no CPU path: a CUDA device (B200, sm_100a) is required")
        self.height, self.width, self.num_scales = height, width, num_scales
        self.device = torch.device(device)
        self._tables = {}
        self.L = _lib.lib()

    def _taps(self, n_in, n_out):
        key = (n_in, n_out)
        if key not in self._tables:
            b, t = lanczos_taps(n_in, n_out)
            self._tables[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(t).to(self.device), t.shape[1])
        return self._tables[key]

    def _p(self, t):
        return ctypes.c_void_p(0 if t is None else t.data_ptr())

    def resize(self, img, height, width, flip=None):
        """uint8 [N,H,W,3] -> [N,height,width,3]: Pillow's two-pass LANCZOS resize (horizontal, then vertical); `flip`
        (uint8 [N]) mirrors the source image first (mono_dataset.py:164 `get_color(..., do_flip)`)."""
        N, H, W, C = img.shape
        img = img.contiguous()
        st = ops._stream()
        if W != width:
            b, t, k = self._taps(W, width)
            out = torch.empty((N, H, width, C), dtype=torch.uint8, device=img.device)
            _lib.check(self.L.mvd_resample_u8(self._p(img), self._p(out), self._p(b), self._p(t), k, N * H, W, width, C, self._p(flip), H, st),
                       "mvd_resample_u8(horizontal)")
            img = out
        elif flip is not None:
            out = torch.empty_like(img)
            _lib.check(self.L.mvd_flip_copy_u8(self._p(img), self._p(out), N, H, W, C, self._p(flip), st), "mvd_flip_copy_u8")
            img = out
        if H != height:
            b, t, k = self._taps(H, height)
            out = torch.empty((N, height, img.shape[2], C), dtype=torch.uint8, device=img.device)
            _lib.check(self.L.mvd_resample_u8(self._p(img), self._p(out), self._p(b), self._p(t), k, N, H, height, img.shape[2] * C,
                                              self._p(None), 0, st), "mvd_resample_u8(vertical)")
            img = out
        ops.launch_counter["n"] += 2
        return img

    def to_tensor(self, img):
        N, H, W, _ = img.shape
        out = torch.empty((N, 3, H, W), dtype=torch.float32, device=img.device)
        _lib.check(self.L.mvd_u8_to_tensor(self._p(img), self._p(out), N, H, W, ops._stream()), "mvd_u8_to_tensor")
        ops.launch_counter["n"] += 1
        return out

    def color_jitter(self, img, orders, factors, active=None):
        """transforms.ColorJitter on a batch of uint8 images [N,H,W,3] (a copy is returned): image n applies the four
        operations in the order `orders[n]` (permutation of 0 brightness, 1 contrast, 2 saturation, 3 hue) with
        `factors[n] = (b, c, s, h)`; images with active[n] == 0 stay unchanged."""
        N, H, W, _ = img.shape
        img = img.clone()
        dev = img.device
        fac = torch.tensor(factors, dtype=torch.float32).reshape(N, 4)
        shift = torch.tensor([int(np.uint8(np.int32(float(f[3]) * 255))) for f in factors], dtype=torch.uint8, device=dev)
        sums = torch.empty(N, dtype=torch.int64, device=dev)
        act = torch.ones(N, dtype=torch.uint8) if active is None else torch.as_tensor(active, dtype=torch.uint8).cpu()
        st = ops._stream()
        for slot in range(4):
            for op in range(4):                           # one launch per (slot, operation) over the images that run it there
                sel = torch.tensor([1 if (int(orders[n][slot]) == op and int(act[n])) else 0 for n in range(N)], dtype=torch.uint8)
                if not bool(sel.any()):
                    continue
                sel_d = sel.to(dev)
                if op < 3:
                    f = fac[:, op].contiguous().to(dev)
                    _lib.check(self.L.mvd_jitter_blend_u8(self._p(img), N, H, W, op, self._p(f), self._p(sums), self._p(sel_d), st),
                               "mvd_jitter_blend_u8")
                else:
                    _lib.check(self.L.mvd_jitter_hue_u8(self._p(img), N, H, W, self._p(shift), self._p(sel_d), st), "mvd_jitter_hue_u8")
                ops.launch_counter["n"] += 1
        return img

    def draw_jitter(self, frame_ids, aug, generator=None):
        """Jitter parameters per item and (frame, scale), drawn item by item in the order `MonoDataset.preprocess` calls its
        `color_aug` (mono_dataset.py:115-126): first the native-resolution image of every frame (results the reference
        throws away, but the draws happen), then scales 0..3 frame by frame."""
        ident = ([0, 1, 2, 3], [1.0, 1.0, 1.0, 0.0])
        out = []
        for n in range(len(aug)):
            d = {}
            if aug[n]:
                for _ in frame_ids:
                    sample_color_jitter(generator)
                for f in frame_ids:
                    for s in range(self.num_scales):
                        d[(f, s)] = sample_color_jitter(generator)
            else:
                for f in frame_ids:
                    for s in range(self.num_scales):
                        d[(f, s)] = ident
            out.append(d)
        return out

    def __call__(self, frames, K_norm, flip=None, color_aug=None, generator=None):
        """frames: {frame_id: uint8 [B,Hn,Wn,3] on the device} (decoded images at native resolution);
        K_norm: normalised 4x4 intrinsics (kitti_dataset.py:26-29); flip / color_aug: per-item booleans [B] (None: nothing).
        With color_aug, every (frame, scale) image of an augmented item draws its own jitter parameters, as the reference's
        `color_aug(f)` does (mono_dataset.py:124); `generator` is the torch CPU generator of those draws.
        Returns the item dict of mono_dataset.py:134-154 on the device."""
        B = next(iter(frames.values())).shape[0]
        flip_d = None if flip is None else torch.as_tensor(flip, dtype=torch.uint8).to(self.device)
        aug = [0] * B if color_aug is None else [int(bool(a)) for a in color_aug]
        params = self.draw_jitter(list(frames), aug, generator) if any(aug) else None
        item = {}
        for f, native in frames.items():
            prev = native
            for s in range(self.num_scales):
                prev = self.resize(prev, self.height // 2 ** s, self.width // 2 ** s, flip_d if s == 0 else None)
                item[("color", f, s)] = self.to_tensor(prev)
                if params is not None:
                    pr = [params[n][(f, s)] for n in range(B)]
                    jit = self.color_jitter(prev, [q[0] for q in pr], [q[1] for q in pr], aug)
                    item[("color_aug", f, s)] = self.to_tensor(jit)
                else:
                    item[("color_aug", f, s)] = item[("color", f, s)].clone()
        Kn = torch.as_tensor(K_norm, dtype=torch.float32)
        for s in range(self.num_scales):
            K = Kn.clone()
            K[0, :] *= self.width // 2 ** s
            K[1, :] *= self.height // 2 ** s
            item[("K", s)] = K.repeat(B, 1, 1).to(self.device)
            item[("inv_K", s)] = torch.linalg.pinv(K).repeat(B, 1, 1).to(self.device)
        return item
