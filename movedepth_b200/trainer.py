"""Trainer with the reference's API surface (movedepth/trainer.py:33-911) on the B200-native hot path.

Same constructor (`Trainer(options)`), same `models` dict keys, same method names
(`train`, `run_epoch`, `process_batch`, `predict_poses`, `compute_losses`, `save_model`, ...), same
item-dict schema for batches (movedepth/datasets/mono_dataset.py:134-154) and the same checkpoint
files.  What differs is how a step executes:

  * the cost volume is built by one fused kernel straight into the layout reg3d consumes
    (no [B,D,C,h,w] tensor, no Python loop over the batch, no hypothesis tensor);
  * softmax + entropy + local-max and the convex upsampling are single kernels;
  * all parameters / gradients / Adam moments live in flat fp32 arenas: one fused Adam kernel per
    parameter group, and data-parallel training all-reduces the gradient arena over NCCL in two
    pieces (cost-volume branch first, overlapping the mono/pose backward) instead of eight DDP
    reducers;
  * the step is two autograd graphs joined only by detached tensors: they are
    back-propagated separately so the first gradient arena can be reduced while the second runs;
  * convolutions run channels-last on cuDNN's tensor-core kernels under the 3xTF32 split policy
    of movedepth_b200/precision.py (fp32 available for bit-level parity runs).

There is no CPU path: `--no_cuda` raises.
"""
import contextlib
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import eventlog, networks, ops
from . import norm as NM
from . import precision as PR
from . import photometric as photo
from .layers import (disp_to_depth, get_smooth_loss, transformation_from_parameters, hypothesis_ratios,
                     fused_group_costvol, convex_upsample_layer, compute_depth_errors)

GROUP0 = ("mono_encoder", "mono_depth", "pose_encoder", "pose", "up")   # lr            (trainer.py:62-141)
GROUP1 = ("mask_cnn", "mvs_encoder", "reg3d")                           # lr * lr_fac


def kitti_intrinsics(batch, height, width, device="cpu"):
    """Normalised KITTI K (movedepth/datasets/kitti_dataset.py:26-29) scaled per
    mono_dataset.py:209-218 -> (K, inv_K) [B,4,4]."""
    K = torch.tensor([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float64)
    K[0] *= width
    K[1] *= height
    return (K.float().repeat(batch, 1, 1).to(device), torch.linalg.pinv(K).float().repeat(batch, 1, 1).to(device))


class SyntheticKITTI:
    """Iterable of synthetic KITTI-shape item dicts in pinned host memory (U[0,1) images, KITTI
    intrinsics at 4 scales), schema of movedepth/datasets/mono_dataset.py:134-154."""

    def __init__(self, opt, batch_size, steps, seed=1, pin=True, smooth=False, shift_px=2):
        """smooth=True: band-limited noise (bicubic-upsampled low-resolution noise) with the source frames shifted by
        `shift_px` pixels per frame index, i.e. images with image-like statistics and a consistent inter-frame motion."""
        self.opt, self.batch_size, self.steps, self.seed, self.pin = opt, batch_size, steps, seed, pin
        self.smooth, self.shift_px = smooth, shift_px

    def __len__(self):
        return self.steps

    def make(self, g):
        o, B = self.opt, self.batch_size
        item = {}
        base = None
        for f in o.frame_ids:
            if self.smooth:
                if base is None:
                    lo = torch.rand(B, 3, o.height // 8, (o.width + 64) // 8, generator=g)
                    base = F.interpolate(lo, size=(o.height, o.width + 64), mode="bicubic", align_corners=False).clamp(0, 1)
                off = 32 + self.shift_px * f
                img = base[:, :, :, off:off + o.width].contiguous()
            else:
                img = torch.rand(B, 3, o.height, o.width, generator=g)
            for s in range(4):
                im = img if s == 0 else F.interpolate(img, size=(o.height // 2 ** s, o.width // 2 ** s), mode="area")
                item[("color", f, s)] = im
                item[("color_aug", f, s)] = im.clone()
        for s in range(4):
            item[("K", s)], item[("inv_K", s)] = kitti_intrinsics(B, o.height // 2 ** s, o.width // 2 ** s)
        if self.pin and torch.cuda.is_available():
            item = {k: v.pin_memory() for k, v in item.items()}
        return item

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.steps):
            yield self.make(g)


def _arena_view(flat, p):
    """A view of `flat` with p's logical shape.  Convolution weights (4-D / 5-D) are stored channels-last (dim 1
    innermost: [Co, k.., Ci]) -- the layout cuDNN's NHWC kernels and their weight gradients use -- so no per-step layout
    conversion of weights or weight gradients is left; every other parameter is stored contiguously."""
    if p.dim() in (4, 5):
        shape = list(p.shape)
        phys = [shape[0]] + shape[2:] + [shape[1]]
        perm = [0, p.dim() - 1] + list(range(1, p.dim() - 1))
        return flat.view(phys).permute(perm)
    return flat.view_as(p)


class FlatArena:
    """All trainable parameters of a group re-homed into one flat fp32 buffer (+ gradient and Adam
    moment buffers of the same shape), so the optimizer and the gradient all-reduce are single
    calls.  Each tensor starts on a 16-byte boundary."""

    def __init__(self, params, device):
        self.params = list(params)
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.data = torch.zeros(n, device=device)
        self.grad = torch.zeros(n, device=device)
        self.exp_avg = torch.zeros(n, device=device)
        self.exp_avg_sq = torch.zeros(n, device=device)
        for p, o in zip(self.params, offs):
            view = _arena_view(self.data[o:o + p.numel()], p)
            view.copy_(p.data)
            p.data = view
            p.grad = _arena_view(self.grad[o:o + p.numel()], p)
        self.offsets = offs
        self._static = None
        # pinned tables for captured collects (allocated here, outside any capture); a graph holds at most a few per arena
        self._pinned_tabs = [torch.zeros(len(self.params), 14, dtype=torch.int64).pin_memory() for _ in range(8)] \
            if self.data.is_cuda else []
        self._pinned_next = 0

    def rebind_grads(self):
        for p, o in zip(self.params, self.offsets):
            p.grad = _arena_view(self.grad[o:o + p.numel()], p)

    def moment_view(self, buf, i):
        """View of a moment / gradient buffer with parameter i's logical shape."""
        p, o = self.params[i], self.offsets[i]
        return _arena_view(buf[o:o + p.numel()], p)

    # ---- one-launch gradient collection (mvd_gather_segments)
    def _gather_static(self):
        if self._static is None:
            chunk = ops._lib.lib().mvd_gather_chunk()
            rows, blocks = [], []
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                phys = list(p.grad.permute(*([0] + list(range(2, p.dim())) + [1])).shape) if p.dim() in (4, 5) else list(p.shape)
                phys = [1] * (5 - len(phys)) + phys if len(phys) <= 5 else None
                assert phys is not None, "parameters with more than 5 dimensions are not supported"
                rows.append((o, p.numel(), phys))
                blocks += [(i, c) for c in range((p.numel() + chunk - 1) // chunk)]
            bm = torch.tensor(blocks, dtype=torch.int32, device=self.data.device)
            self._static = (rows, bm, len(blocks))
        return self._static

    def gather_table(self, grads):
        """Host int64 [nseg,14] segment table for `grads` (one tensor or None per parameter)."""
        rows, _, _ = self._gather_static()
        tab = torch.zeros(len(rows), 14, dtype=torch.int64)
        for i, (g, p, (o, n, phys)) in enumerate(zip(grads, self.params, rows)):
            tab[i, 1] = o
            tab[i, 4:9] = torch.tensor(phys)
            if g is None:               # numel 0: the segment is left untouched (zero from the start if it never gets a gradient)
                continue
            tab[i, 2] = n
            assert g.shape == p.shape and g.dtype == torch.float32, (g.shape, p.shape, g.dtype)
            gs = g.permute(*([0] + list(range(2, g.dim())) + [1])) if g.dim() in (4, 5) else g     # destination physical order
            st = [0] * (5 - gs.dim()) + list(gs.stride())
            tab[i, 0] = g.data_ptr()
            tab[i, 3] = 1 if gs.is_contiguous() else 0
            tab[i, 9:14] = torch.tensor(st)
        return tab

    def collect(self, grads, pinned=False):
        """Write the gradient tensors returned by torch.autograd.grad into the flat gradient buffer with one kernel
        (no zero_grad, no per-parameter accumulate); parameters whose entry is None keep what the buffer holds.  `grads`
        must stay alive until the kernel has run; returns the objects the caller has to keep alive (CUDA-graph capture:
        the pinned table is re-read at every replay)."""
        _, bm, nblocks = self._gather_static()
        tab = self.gather_table(grads)
        if pinned:                  # capture: the copy node re-reads a persistent pinned table at every replay
            slot = self._pinned_tabs[self._pinned_next % len(self._pinned_tabs)]
            self._pinned_next += 1
            slot.copy_(tab)
            tab = slot
        dev_tab = tab.to(self.data.device, non_blocking=pinned)
        rc = ops._lib.lib().mvd_gather_segments(ops._p(dev_tab), ops._p(bm), nblocks, ops._p(self.grad), ops._stream())
        ops._lib.check(rc, "mvd_gather_segments")
        ops.launch_counter["n"] += 1
        return tab, dev_tab, list(grads)


class Trainer:
    def __init__(self, options, train_loader=None, val_loader=None):
        self.opt = o = options
        self.log_path = os.path.join(o.log_dir, o.model_name)
        self.writers = {}                      # mode -> eventlog.SummaryWriter (movedepth/trainer.py:147-151), opened on first use
        assert o.height % 32 == 0, "'height' must be a multiple of 32"
        assert o.width % 32 == 0, "'width' must be a multiple of 32"
        assert o.frame_ids[0] == 0, "frame_ids must start with 0"
        assert len(o.frame_ids) > 1, "frame_ids must have more than 1 frame specified"
        if o.no_cuda or not torch.cuda.is_available():
            raise RuntimeError("movedepth_b200 has no CPU path: a CUDA device (B200, sm_100a) is required")
        ops._lib.lib()                                    # fail loudly now if the extension is missing

        self.local_rank = int(os.environ.get("LOCAL_RANK", o.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.world_size, self.rank = 1, 0
        if o.ddp:
            if not dist.is_initialized():
                dist.init_process_group(backend="nccl")
            self.world_size, self.rank = dist.get_world_size(), dist.get_rank()

        self.num_scales = len(o.scales)
        self.num_input_frames = len(o.frame_ids)
        self.num_pose_frames = 2
        self.matching_ids = o.matching_ids
        self.precision = getattr(o, "b200_conv_precision", "3xtf32")
        PR.set_policy(self.precision, split_backward=getattr(o, "b200_split_backward", False))
        torch.backends.cudnn.benchmark = True
        # how many cuDNN algorithms the autotuner tries per conv shape (0 = all): trying all of them finds faster tensor-core
        # gradient kernels (-2.5 % step time); the fp32 policy (bit-level parity runs) keeps PyTorch's default of 10
        limit = os.environ.get("MVD_CUDNN_BENCHMARK_LIMIT")
        torch.backends.cudnn.benchmark_limit = int(limit) if limit else (10 if self.precision == "fp32" else 0)

        # ---- sub-models (trainer.py:65-131)
        pretrained = o.weights_init == "pretrained"
        m = {}
        m["mono_encoder"] = networks.ResnetEncoder(o.res_arch, pretrained)
        m["mono_depth"] = networks.DepthDecoder(m["mono_encoder"].num_ch_enc, o.scales)
        m["pose_encoder"] = networks.ResnetEncoder(o.res_arch, pretrained, num_input_images=self.num_pose_frames)
        m["pose"] = networks.PoseDecoder(m["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
        m["mask_cnn"] = networks.UncertNet()
        m["mvs_encoder"] = networks.FPN4(base_channels=8, scale=o.prior_scale, dcn=o.dcn)
        if o.num_depth_bins >= 8:
            m["reg3d"] = networks.reg3d(o.reg3d_c, o.reg3d_c, 3)
        else:
            m["reg3d"] = networks.reg2d(o.reg3d_c, o.reg3d_c)      # same constructor arguments as evaluate_depth.build_models
        if o.convex_up:
            m["up"] = convex_upsample_layer(8 * 2 ** o.prior_scale, o.prior_scale)
        for k in m:
            if o.ddp and self.world_size > 1:
                m[k] = nn.SyncBatchNorm.convert_sync_batchnorm(m[k])
            PR.adopt(m[k])                                # torchvision convs follow the precision policy too
            m[k].to(self.device)
        self.models = m
        self.parameters_to_train = [p for k in GROUP0 if k in m for p in m[k].parameters()]
        self.mvs_parameters_to_train = [p for k in GROUP1 if k in m for p in m[k].parameters()]

        if o.mask_mvs_geo:
            # the reference reads outputs[("geo_mask", f)] (trainer.py:654-656) but nothing ever writes it: KeyError there
            raise NotImplementedError("--mask_mvs_geo: the reference never produces the geo_mask it multiplies in "
                                      "(movedepth/trainer.py:654-656 raises KeyError); not supported")

        # ---- flat arenas + fused Adam (replaces optim.Adam + StepLR, trainer.py:137-141)
        self.arenas = [FlatArena(self.parameters_to_train, self.device),
                       FlatArena(self.mvs_parameters_to_train, self.device)]
        self.base_lrs = [o.learning_rate, o.learning_rate * o.lr_fac]
        self.opt_step = 0
        self.epoch = 0
        self.step = 0

        # checkpoints are loaded into the arenas (parameters are views of them), Adam moments included
        if o.load_weights_folder is not None:
            self.load_model()
        if o.mono_weights_folder is not None:
            self.load_mono_model()
        if o.ddp and self.world_size > 1:                 # what the DDP constructor's broadcast does
            for a in self.arenas:
                dist.broadcast(a.data, 0)
            for k in m:
                for t in m[k].buffers():
                    dist.broadcast(t.data, 0)

        # ---- data: any iterable of item dicts; synthetic KITTI-shape tensors by default
        steps = 100
        self.train_loader = train_loader if train_loader is not None else SyntheticKITTI(o, o.batch_size, steps, seed=1 + self.rank)
        self.val_loader = val_loader if val_loader is not None else SyntheticKITTI(o, o.batch_size, 4, seed=10_000 + self.rank)
        self.val_iter = iter(self.val_loader)
        self.depth_metric_names = ["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"]
        if o.ddp:
            o.log_frequency = max(1, o.log_frequency // self.world_size)
        self._aug_box = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._static_inputs, self._graphs, self._graph_warm, self._graph_pool, self._graph_stream = None, {}, {}, None, None
        # auto-mask tie-break noise (trainer.py:600,641,698 draw it on the CPU and copy it over): a device generator of this
        # trainer fills static buffers OUTSIDE the captured graph, so eager and graph steps with the same seed see the same
        # noise and a replay never re-uses the noise of the capture
        if getattr(o, "b200_cudnn_benchmark", False) or os.environ.get("MVD_CUDNN_BENCHMARK") == "1":
            torch.backends.cudnn.benchmark = True
        self._two_streams = not getattr(o, "b200_one_stream", False)
        self._mvs_stream = torch.cuda.Stream() if self._two_streams else None
        if self.device.type == "cuda":
            for chan in (0, 1):                            # the one-kernel BatchNorm's workspaces exist before any graph capture
                NM.workspace(self.device, chan)
        self.noise_generator = torch.Generator(device=self.device)
        self.noise_generator.manual_seed(torch.initial_seed() + self.rank)
        self._noise_buf = None
        if self.rank == 0:
            self.save_opts()
        self.set_train()

    # ------------------------------------------------------------------ mode switches
    def set_train(self):
        for mod in self.models.values():
            mod.train()

    def set_eval(self):
        for mod in self.models.values():
            mod.eval()

    def current_lrs(self):
        decay = 0.1 ** (self.epoch // self.opt.scheduler_step_size)
        return [lr * decay for lr in self.base_lrs]

    def _tf32(self, branch=None):
        """Apply the conv arithmetic policy (movedepth_b200/precision.py): fp32 | 3xtf32 | tf32."""
        PR.set_policy(self.precision)

    # ------------------------------------------------------------------ training loop
    def train(self):
        """Run the entire training pipeline (movedepth/trainer.py:244-256)."""
        self.epoch, self.step = 0, 0
        self.start_time = time.time()
        for self.epoch in range(self.opt.num_epochs):
            self.run_epoch()
            if (self.epoch + 1) % self.opt.save_frequency == 0 and self.epoch > 15:
                self.save_model()

    def run_epoch(self):
        """One pass over the train loader (movedepth/trainer.py:258-295)."""
        self.set_train()
        for batch_idx, inputs in enumerate(self.train_loader):
            t0 = time.time()
            outputs, losses = self.train_step(inputs)
            early = batch_idx % self.opt.log_frequency == 0 and self.step < 2000
            late = self.step % 2000 == 0
            if early or late:
                if self.rank == 0:
                    self.log_time(batch_idx, time.time() - t0, float(losses["loss"]))
                if "depth_gt" in inputs:
                    self.compute_depth_losses(inputs, outputs, losses)
                if self.rank == 0:
                    self.log("train", inputs, outputs, losses)
                self.val()
            if self.opt.save_intermediate_models and late:
                self.save_model(save_step=True)
            self.step += 1

    def train_step(self, inputs, noise=None, mask_xy=None):
        """process_batch + zero_grad + backward + optimizer.step (movedepth/trainer.py:269-272).
        With `--b200_cuda_graph` the whole forward + backward + gradient exchange is one CUDA graph replayed per step:
        `inputs` (and the `noise` / `mask_xy` overrides of the parity tests, if given) are copied into static device
        buffers first."""
        if getattr(self.opt, "b200_cuda_graph", False):
            return self._graph_step(inputs, noise, mask_xy)
        outputs, losses = self._forward_backward(inputs, noise, mask_xy)
        self._optimizer_step()
        return outputs, losses

    def _forward_backward(self, inputs, noise=None, mask_xy=None):
        outputs, losses = self.process_batch(inputs, is_train=True, noise=noise, mask_xy=mask_xy)
        multi = self.opt.ddp and self.world_size > 1
        capturing = torch.cuda.is_current_stream_capturing()
        # The two graphs share nothing but detached tensors: they are back-propagated separately.
        # torch.autograd.grad hands back the raw gradient tensors (no zero_grad, no per-parameter accumulate kernels); one
        # gather kernel per parameter group writes them into the flat arena the all-reduce and the fused Adam kernel work on.
        # Data-parallel training: the mono / pose arena (27 M parameters, 107 MB) is all-reduced as soon as its (shorter)
        # back-propagation is done, overlapping the rest of the cost-volume backward; the cost-volume arena (1.4 M parameters)
        # and the `up` tail of arena 0, whose gradients come from the cost-volume graph, are two small calls.
        a0, a1 = self.arenas
        p0, p1 = a0.params, a1.params
        main = torch.cuda.current_stream()
        side = self._mvs_stream if self._two_streams else None
        keep = []
        # cost-volume graph: issued first (host order) on its own stream; the mono / pose graph follows on the main stream,
        # so the two back-propagations run concurrently on the device
        if side is not None:
            side.wait_stream(main)
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            self._tf32("mvs")
            g = torch.autograd.grad(losses["_mvs_total"], p1 + p0, allow_unused=True)
            g1, g0_mvs = g[:len(p1)], g[len(p1):]
            keep.append(a1.collect(g1, pinned=capturing))
            keep.append(a0.collect(g0_mvs, pinned=capturing))         # the `up` head: arena 0's tail, nothing else
            tail = self._arena0_tail(g0_mvs)
            if multi:
                dist.all_reduce(a1.grad)
                if tail < a0.numel:
                    dist.all_reduce(a0.grad[tail:])
        self._tf32("mono")
        g0 = torch.autograd.grad(losses["_mono_total"], p0, allow_unused=True)
        assert all(x is None or y is None for x, y in zip(g0, g0_mvs)), "a parameter receives gradients from both graphs"
        keep.append(a0.collect(g0, pinned=capturing))
        if multi:
            dist.all_reduce(a0.grad[:tail])          # 107 MB: overlaps whatever is left of the cost-volume backward
        if side is not None:
            main.wait_stream(side)
        outputs["_keepalive"] = keep
        return outputs, losses

    def _arena0_tail(self, g0_mvs):
        """Offset of the first arena-0 parameter that gets its gradient from the cost-volume graph (the `up` head sits at
        the end of the arena); arena 0's numel when there is none."""
        a0 = self.arenas[0]
        first = min((i for i, g in enumerate(g0_mvs) if g is not None), default=None)
        return a0.offsets[first] if first is not None else a0.numel

    def _optimizer_step(self):
        self.opt_step += 1
        for a, lr in zip(self.arenas, self.current_lrs()):
            ops.adam_step(a.data, a.grad, a.exp_avg, a.exp_avg_sq, self.opt_step, lr, grad_scale=1.0 / self.world_size)

    # ------------------------------------------------------------------ random draws of a step
    def _num_noise_maps(self):
        o = self.opt
        return (0 if o.disable_automasking else len(o.scales)) + (2 if o.mask_mvs_auto else 0)

    def _draw_noise(self, B, given=None):
        """The step's N(0,1) tie-break maps as a list of [B,1,H,W] views of one static buffer: drawn from this trainer's
        device generator, or copied from `given` (parity tests)."""
        o, n = self.opt, self._num_noise_maps()
        if self._noise_buf is None or self._noise_buf.shape[1] != B:
            self._noise_buf = torch.zeros(max(n, 1), B, 1, o.height, o.width, device=self.device)
        if given is not None:
            for i, t in enumerate(list(given)[:n]):
                self._noise_buf[i].copy_(t, non_blocking=True)
        elif n:
            self._noise_buf.normal_(generator=self.noise_generator)
        return [self._noise_buf[i] for i in range(n)]

    def _set_aug_box(self, mask_xy=None):
        """Corner of the masked-augmentation box (layers.py:64-65: np.random.randint, x before y) -> device."""
        o = self.opt
        fh, fw = o.height // 3, o.width // 3
        if mask_xy is None:
            mask_xy = (np.random.randint(0, o.width - fw), np.random.randint(0, o.height - fh))
        # pageable source: the copy is staged before the call returns, so the host may run ahead safely
        self._aug_box.copy_(torch.tensor([int(mask_xy[0]), int(mask_xy[1])], dtype=torch.int64))

    # ------------------------------------------------------------------ CUDA-graph step
    GRAPH_WARMUP = 3          # eager steps before capture (cuDNN autotuning, lazy initialisation, NCCL warm-up)

    def _stage_inputs(self, inputs, noise=None, mask_xy=None):
        """host (pinned) or device item dict -> the static device buffers the graph reads; the step's random draws
        (tie-break noise, augmentation box) are made here, outside the graph."""
        if self._static_inputs is None:
            self._static_inputs = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device)
                                   for k, v in inputs.items() if torch.is_tensor(v)}
        for k, buf in self._static_inputs.items():
            buf.copy_(inputs[k], non_blocking=True)
        B = self._static_inputs[("color_aug", 0, 0)].shape[0]
        self._draw_noise(B, noise)
        self._set_aug_box(mask_xy)
        return dict(self._static_inputs)

    def _graph_step(self, inputs, noise=None, mask_xy=None):
        static = self._stage_inputs(inputs, noise, mask_xy)
        key = self.epoch > self.opt.ztrans_start_epc      # the only data-independent branch of the step (trainer.py:336)
        g = self._graphs.get(key)
        if g is None:
            # Warm-up and capture both run on one side stream: autograd's AccumulateGrad nodes remember the stream they
            # were created on, and a node created on the legacy default stream would invalidate the capture.
            if self._graph_stream is None:
                self._graph_stream = torch.cuda.Stream()
            side, cur = self._graph_stream, torch.cuda.current_stream()
            side.wait_stream(cur)
            if self._graph_warm.get(key, 0) < self.GRAPH_WARMUP:
                self._graph_warm[key] = self._graph_warm.get(key, 0) + 1
                with torch.cuda.stream(side):
                    outputs, losses = self._forward_backward(static, noise="preset", mask_xy="preset")
                    self._optimizer_step()
                cur.wait_stream(side)
                return outputs, losses
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_counter["n"]
            with torch.cuda.graph(graph, pool=self._graph_pool, stream=side):
                outputs, losses = self._forward_backward(static, noise="preset", mask_xy="preset")
            g = self._graphs[key] = (graph, outputs, losses, ops.launch_counter["n"] - n0)
            if self._graph_pool is None:
                self._graph_pool = graph.pool()
        graph, outputs, losses, launches = g
        graph.replay()
        ops.launch_counter["n"] += launches
        self._optimizer_step()
        return outputs, losses

    # ------------------------------------------------------------------ forward + losses
    def predict_poses(self, inputs):
        """Pose net on (earlier, later) frame pairs (movedepth/trainer.py:445-468)."""
        outputs = {}
        for f in self.opt.frame_ids[1:]:
            if f == "s":
                continue
            a, b = inputs[("color_aug", f, 0)], inputs[("color_aug", 0, 0)]
            pair = [a, b] if f < 0 else [b, a]
            feats = [self.models["pose_encoder"](torch.cat(pair, 1).contiguous(memory_format=torch.channels_last))]
            axisangle, translation = self.models["pose"](feats)
            outputs[("axisangle", 0, f)], outputs[("translation", 0, f)] = axisangle, translation
            outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(axisangle[:, 0], translation[:, 0], invert=(f < 0))
        for f in self.matching_ids[1:]:
            inputs[("relative_pose", f)] = outputs[("cam_T_cam", 0, f)].clone().detach()
        return outputs

    def _volume_logits(self, ref_feat, src_feats, inputs, prior, ratio, poses):
        """cost volume over the matching frames -> reg3d logits (trainer.py:349-366)."""
        K, invK = inputs[("K", 2)], inputs[("inv_K", 2)]       # scale index 2 is hard-coded in the reference
        # channels-last-3d volume ([B,D,h,w,G] in memory): what cuDNN's tensor-core NDHWC kernels consume
        vols = [fused_group_costvol(ref_feat, src_feats[i], K, invK, poses[:, i], prior, ratio, self.opt.reg3d_c,
                                    layout=ops.LAYOUT_BDHWG) for i in range(len(src_feats))]
        if len(vols) == 1:
            vol = vols[0]        # single view: the view weight w/(1e-8+w) is 1 to within 2e-7
        else:
            wsum, acc = 1e-8, 0
            for v in vols:       # [B,G,D,h,w]: mean over D, softmax over G, max (trainer.py:360)
                wgt = torch.softmax(v.mean(2), dim=1).max(1)[0]
                wsum = wsum + wgt
                acc = acc + wgt[:, None, None] * v
            vol = acc / wsum[:, None, None]
        return self.models["reg3d"].forward_volume(vol), vol

    def process_batch(self, inputs, is_train=False, noise=None, mask_xy=None):
        """Forward pass and all losses for one minibatch (movedepth/trainer.py:297-442).
        `noise` (list of [B,1,H,W] N(0,1) tensors) and `mask_xy` (box corner) override the random
        draws so that parity tests can reproduce the reference bit for bit.

        The step is two autograd graphs joined only by detached tensors: the mono / pose graph and the
        cost-volume graph.  With `--b200_two_streams` (default) they are issued on two CUDA streams -- the matching-feature
        network starts at once, the cost-volume part waits for the mono prior and the poses -- so that the small kernels of
        one graph fill the tails of the other and, in data-parallel training, one graph's SyncBatchNorm exchange latency
        hides behind the other's compute.  Captured in the step's CUDA graph the two streams become parallel branches."""
        o = self.opt
        inputs = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in inputs.items()}
        B = inputs[("color_aug", 0, 0)].shape[0]
        if isinstance(noise, str):                       # "preset": the graph step already filled the static noise buffers
            noise = [self._noise_buf[i] for i in range(self._num_noise_maps())]
        elif is_train or noise is not None:
            noise = self._draw_noise(B, noise)
        fh, fw = o.height // 3, o.width // 3
        if not isinstance(mask_xy, str):     # "preset": the caller already wrote the box corner into self._aug_box (graph step)
            self._set_aug_box(mask_xy)

        main = torch.cuda.current_stream()
        side = self._mvs_stream if self._two_streams else None
        # Cross-stream discipline: a tensor that the OTHER stream's graph saves for its backward must not be allocated on this
        # stream -- when autograd releases it the caching allocator would hand the block to the next allocation of its own
        # stream while the other stream's kernel may still be pending (inside a captured graph: a genuine race between the two
        # branches).  So everything the cost-volume graph keeps is either allocated on its stream (copies below) or outlives
        # the step (`outputs`, static inputs); freshly staged input copies are marked as used by both streams.
        if side is not None:
            for v in inputs.values():
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(side)

        def on_mvs():
            return torch.cuda.stream(side) if side is not None else contextlib.nullcontext()

        # ---------------- cost-volume graph, part 1: matching features (independent of the mono / pose graph)
        if side is not None:
            side.wait_stream(main)
        with on_mvs():
            NM.channel = 1
            self._tf32("mvs")
            enc = self.models["mvs_encoder"]
            cl = torch.channels_last   # NHWC through FPN4: its output is then already the layout K1's TMA boxes read
            ref_img = inputs[("color_aug", 0, 0)].contiguous(memory_format=cl)
            ref_feat, ref_ctx = enc(ref_img)
            src_feats = [enc(inputs[("color_aug", f, 0)].contiguous(memory_format=cl))[0] for f in self.matching_ids[1:]]
            # masked augmentation (trainer.py:374-376): zero a random box of the reference image
            ys = torch.arange(o.height, device=self.device).view(1, 1, -1, 1)
            xs = torch.arange(o.width, device=self.device).view(1, 1, 1, -1)
            inside = (xs >= self._aug_box[0]) & (xs < self._aug_box[0] + fw) & (ys >= self._aug_box[1]) & (ys < self._aug_box[1] + fh)
            aug_mask = (~inside).float().expand(B, 3, o.height, o.width)
            aug_feat, _ = enc((ref_img * aug_mask).contiguous(memory_format=cl))

        # ---------------- mono / pose graph
        NM.channel = 0
        self._tf32("mono")
        outputs = self.predict_poses(inputs)
        poses = torch.stack([inputs[("relative_pose", f)] for f in self.matching_ids[1:]], 1)   # [B,M,4,4]
        ref_cl = inputs[("color_aug", 0, 0)].contiguous(memory_format=torch.channels_last)    # NHWC for cuDNN's tensor-core kernels
        outputs.update(self.models["mono_depth"](self.models["mono_encoder"](ref_cl)))
        disp_prior = outputs[("disp", o.prior_scale)].detach()          # views of tensors `outputs` keeps alive
        disp_full = outputs[("disp", 0)].detach()
        if side is not None:
            prior_ready = torch.cuda.Event()
            prior_ready.record(main)
        losses = self.compute_losses(inputs, outputs, is_mvs=False, noise=noise)

        # ---------------- cost-volume graph, part 2
        with on_mvs():
            if side is not None:
                side.wait_event(prior_ready)
            NM.channel = 1
            self._tf32("mvs")
            poses = poses.clone()                                       # this stream's own copies (see above)
            _, mono_depth = disp_to_depth(disp_full, o.min_depth, o.max_depth)
            # hypotheses around the mono prior (trainer.py:333-346), separable form
            prior = 1.0 / (1.0 / o.max_depth + disp_prior * (1.0 / o.min_depth - 1.0 / o.max_depth))
            if self.epoch > o.ztrans_start_epc:
                s = o.depth_bin_fac * o.z_scale * poses[:, 0, 2, 3]
            else:
                s = torch.full((B,), float(o.depth_bin_fac), device=self.device)
            ratio = hypothesis_ratios(o.num_depth_bins, s, self.device, o.schedule_type)          # [B,D]
            inv_a = 1.0 / (prior[:, 0] * ratio[:, -1].view(B, 1, 1))
            inv_b = 1.0 / (prior[:, 0] * ratio[:, 0].view(B, 1, 1))
            outputs["depth_prior"], outputs["hypothesis_ratio"] = prior, ratio

            logits, vol = self._volume_logits(ref_feat, src_feats, inputs, prior, ratio, poses)
            cost_prob, ent, depth_mvs = ops.regress_depth(logits, inv_a, inv_b, o.norm_radius, want_prob=bool(o.mask_mvs_conf))
            trust = self.models["mask_cnn"](ent)
            outputs["cost_volume"], outputs["cost_logits"], outputs["depth_mvs_lowres"] = vol, logits, depth_mvs

            # masked-augmentation consistency (trainer.py:377-403)
            logits_aug, _ = self._volume_logits(aug_feat, src_feats, inputs, prior, ratio, poses)
            _, _, depth_aug = ops.regress_depth(logits_aug, inv_a, inv_b, o.norm_radius)
            # mean smooth-L1 over the pixels the resized box mask selects, one kernel; the weight is applied twice in the
            # reference (trainer.py:399-400)
            masked = ops.masked_smooth_l1(depth_aug, depth_mvs, self._aug_box, o.height, o.width, fh, fw, o.mask_lw * o.mask_lw)
            outputs["masked_depth"], outputs["masked_aug"] = depth_aug, aug_mask

            # upsample + fuse (trainer.py:405-416)
            if o.convex_up:
                depth_up = self.models["up"](depth_mvs, ref_ctx)
            else:
                depth_up = F.interpolate(depth_mvs.unsqueeze(1), [o.height, o.width], mode="bilinear", align_corners=True)[:, 0]
            outputs["depth_mvs"] = depth_up
            trust = F.interpolate(trust, [o.height, o.width], mode="bilinear", align_corners=True)
            outputs["trust_mono_mask"] = trust
            outputs["fused_depth"] = (1 - trust) * depth_up.unsqueeze(1).detach() + trust * mono_depth
            fuse_loss = self.compute_fuse_losses(inputs, outputs, noise=noise)
            if o.mask_mvs_conf:                  # trainer.py:419-422: trilinear-upsampled probability peak above photo_conf
                up = F.interpolate(cost_prob.unsqueeze(1), [o.num_depth_bins, o.height, o.width], mode="trilinear", align_corners=True)
                outputs["photo_conf_map"] = up.max(2)[0] > o.photo_conf
            if o.mask_mvs_dist:                  # trainer.py:423-425
                outputs["dist_mask"] = outputs[("disp", 0)] > o.dist_thres
            mvs_losses = self.compute_losses(inputs, outputs, is_mvs=True, noise=noise)
            mvs_total = mvs_losses["loss"] + masked + fuse_loss
        NM.channel = 0
        if side is not None:
            main.wait_stream(side)

        # ---------------- totals (trainer.py:429-440): loss = mvs + (mono + masked) + fuse
        losses["masked_loss"] = masked
        losses["fuse_reproj_loss"] = fuse_loss
        losses.update({k: v for k, v in mvs_losses.items() if k != "loss"})
        losses["_mono_total"] = losses["loss"]
        losses["_mvs_total"] = mvs_total
        losses["loss"] = losses["_mono_total"] + losses["_mvs_total"]
        return outputs, losses

    def generate_images_pred(self, inputs, outputs, is_mvs=False):
        """Warped source images into `outputs` (movedepth/trainer.py:491-532).  `compute_losses` produces the same images
        as a by-product of the fused warp + SSIM kernel; this entry point is for callers that only want the images:
        `("depth", 0, s)`, `("color", f, s)`, `("color_identity", f, s)`, or `("mvs_color", f)` with is_mvs."""
        o = self.opt
        tgt = inputs[("color", 0, 0)]
        K, invK = inputs[("K", 0)], inputs[("inv_K", 0)]
        with torch.no_grad():
            if is_mvs:
                for f in o.frame_ids[1:]:
                    outputs[("mvs_color", f)] = photo.reprojection_loss(outputs["depth_mvs"], inputs[("color", f, 0)], tgt, K, invK,
                                                                        outputs[("cam_T_cam", 0, f)].detach(), 0)[1]
                return
            for s in o.scales:
                depth = ops.disp_to_depth_full(outputs[("disp", s)].detach(), o.height, o.width, o.min_depth, o.max_depth)
                outputs[("depth", 0, s)] = depth
                for f in o.frame_ids[1:]:
                    outputs[("color", f, s)] = photo.reprojection_loss(depth, inputs[("color", f, 0)], tgt, K, invK,
                                                                       outputs[("cam_T_cam", 0, f)].detach(), 0)[1]
                    outputs[("color_identity", f, s)] = inputs[("color", f, 0)]

    @staticmethod
    def compute_loss_masks(reprojection_loss, identity_reprojection_loss):
        """argmin over [reprojection, identity] == 0 (movedepth/trainer.py:552-567)."""
        if identity_reprojection_loss is None:
            return torch.ones_like(reprojection_loss)
        return (reprojection_loss <= identity_reprojection_loss).float()

    def compute_reprojection_loss(self, pred, target, ssim_lw=None):
        """movedepth/trainer.py:535-550."""
        w = self.opt.ssim_lw if ssim_lw is None else ssim_lw
        return photo.photo_error(pred, target, 0 if self.opt.no_ssim else w)

    def _identity_min(self, inputs, ssim_w):
        tgt = inputs[("color", 0, 0)]
        ident = [photo.identity_loss(inputs[("color", f, 0)], tgt, ssim_w) for f in self.opt.frame_ids[1:]]
        return torch.cat(ident, 1).min(1, keepdim=True)[0]

    def compute_losses(self, inputs, outputs, is_mvs=False, noise=None):
        """Photometric + smoothness losses (movedepth/trainer.py:491-532 warps, 614-724 losses)."""
        o = self.opt
        tgt = inputs[("color", 0, 0)]
        K, invK = inputs[("K", 0)], inputs[("inv_K", 0)]
        ssim_w = 0 if o.no_ssim else o.ssim_lw
        losses = {}
        if is_mvs:
            depth = outputs["depth_mvs"]
            per_src = []
            for f in o.frame_ids[1:]:
                T = outputs[("cam_T_cam", 0, f)].detach()
                l, warped = photo.reprojection_loss(depth, inputs[("color", f, 0)], tgt, K, invK, T, ssim_w)
                outputs[("mvs_color", f)] = warped
                per_src.append(l)
            if o.mask_mvs_auto and noise is not None and len(noise):
                noise.pop(0)                         # drawn by the reference, the mask is then overwritten by ones
            extra = [outputs[k].float() for k, on in (("photo_conf_map", o.mask_mvs_conf), ("dist_mask", o.mask_mvs_dist)) if on]
            if len(per_src) <= 2 and not extra:      # min over sources + mean in one kernel (mask = ones)
                loss, reproj = ops.reproj_select(per_src)
                outputs["reprojection_loss_mask"] = torch.ones_like(reproj)
            else:                                    # optional confidence / distance masks (trainer.py:648-660)
                reproj = torch.cat(per_src, 1).min(1, keepdim=True)[0]
                mask = torch.ones_like(reproj)
                for e in extra:
                    mask = mask * e
                outputs["reprojection_loss_mask"] = mask
                loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
            outputs["mvs_reprojection_loss"] = reproj
            if o.mvs_smooth_loss:
                sm = ops.smooth_loss(depth.unsqueeze(1), tgt, normalize=True)
                losses["mvs_smooth_loss/0"] = sm
                loss = loss + o.disparity_smoothness * sm
            losses["mvs_reproj_loss"] = loss
            losses["loss"] = loss
            return losses

        ident = None if o.disable_automasking else self._identity_min(inputs, ssim_w)
        total = 0
        for s in o.scales:
            disp = outputs[("disp", s)]
            depth = ops.disp_to_depth_full(disp, o.height, o.width, o.min_depth, o.max_depth)   # upsample + disp_to_depth
            outputs[("depth", 0, s)] = depth
            per_src = []
            for f in o.frame_ids[1:]:
                l, warped = photo.reprojection_loss(depth, inputs[("color", f, 0)], tgt, K, invK,
                                                    outputs[("cam_T_cam", 0, f)], ssim_w)
                outputs[("color", f, s)] = warped
                per_src.append(l)
            nz = None
            if ident is not None:
                nz = noise.pop(0).to(self.device) if noise is not None else torch.randn_like(ident)
            if len(per_src) <= 2:                    # min over sources, auto-mask and masked mean in one kernel
                loss, reproj = ops.reproj_select(per_src, ident, nz)
            else:
                reproj = torch.cat(per_src, 1).min(1, keepdim=True)[0]
                # argmin over [reproj, identity] == 0
                mask = (reproj <= ident + nz * 1e-5).float() if ident is not None else torch.ones_like(reproj)
                loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
            if s == 0:
                outputs["mono_reproj_loss"] = reproj
            sm = ops.smooth_loss(disp, inputs[("color", 0, s)], normalize=True)     # mean normalisation + smoothness
            losses["mono_smooth_loss/{}".format(s)] = sm
            loss = loss + o.disparity_smoothness * sm / (2 ** s)
            total = total + loss
            losses["loss/{}".format(s)] = loss
        losses["loss"] = total / self.num_scales
        return losses

    def compute_fuse_losses(self, inputs, outputs, noise=None):
        """L1-only reprojection of the fused depth; trains the trust mask (trainer.py:569-612)."""
        o = self.opt
        tgt = inputs[("color", 0, 0)]
        K, invK = inputs[("K", 0)], inputs[("inv_K", 0)]
        per_src = []
        for f in o.frame_ids[1:]:
            l, warped = photo.reprojection_loss(outputs["fused_depth"], inputs[("color", f, 0)], tgt, K, invK,
                                                outputs[("cam_T_cam", 0, f)].detach(), 0)
            outputs[("mvs_color_fuse", f)] = warped
            per_src.append(l)
        ident = nz = None
        if o.mask_mvs_auto:
            ident = self._identity_min(inputs, 0)
            nz = noise.pop(0).to(self.device) if noise is not None else torch.randn_like(ident)
        if len(per_src) <= 2:
            return ops.reproj_select(per_src, ident, nz)[0]
        reproj = torch.cat(per_src, 1).min(1, keepdim=True)[0]
        mask = (reproj <= ident + nz * 1e-5).float() if ident is not None else torch.ones_like(reproj)
        return (reproj * mask).sum() / (mask.sum() + 1e-7)

    # ------------------------------------------------------------------ validation / logging / checkpoints
    def val(self):
        """One validation minibatch (movedepth/trainer.py:470-489)."""
        self.set_eval()
        try:
            inputs = next(self.val_iter)
        except StopIteration:
            self.val_iter = iter(self.val_loader)
            inputs = next(self.val_iter)
        with torch.no_grad():
            outputs, losses = self.process_batch(inputs)
            if "depth_gt" in inputs:
                self.compute_depth_losses(inputs, outputs, losses)
            if self.rank == 0:
                self.log("val", inputs, outputs, losses)
        self.set_train()
        return losses

    def compute_depth_losses(self, inputs, outputs, losses):
        """KITTI depth metrics on a minibatch, for monitoring (movedepth/trainer.py:726-757)."""
        pred = torch.clamp(F.interpolate(outputs[("depth", 0, 0)], [375, 1242], mode="bilinear", align_corners=False),
                           1e-3, 80).detach()
        gt = inputs["depth_gt"].to(self.device)
        mask = gt > 0
        crop = torch.zeros_like(mask)
        crop[:, :, 153:371, 44:1197] = 1
        mask = mask * crop
        gt, pred = gt[mask], pred[mask]
        pred = torch.clamp(pred * torch.median(gt) / torch.median(pred), min=1e-3, max=80)
        for name, v in zip(self.depth_metric_names, compute_depth_errors(gt, pred)):
            losses[name] = np.array(v.cpu())

    def log_time(self, batch_idx, duration, loss):
        """movedepth/trainer.py:759-770."""
        rate = self.opt.batch_size / max(duration, 1e-9)
        spent = time.time() - getattr(self, "start_time", time.time())
        print("epoch {:>3} | batch {:>6} | examples/s: {:5.1f} | loss: {:.5f} | time elapsed: {:.0f}s".format(
            self.epoch, batch_idx, rate, loss, spent))

    def log(self, mode, inputs, outputs, losses):
        """movedepth/trainer.py:772-793: scalars per loss plus the colour frames, warped predictions and colour-mapped
        disparities of up to four samples, written as a tensorboard events file under <log_path>/<mode>/ by
        `eventlog.SummaryWriter` (this image has no tensorboardX); the scalars also go to scalars.jsonl."""
        NM.check_workspaces()
        d = os.path.join(self.log_path, mode)
        if mode not in self.writers:
            self.writers[mode] = eventlog.SummaryWriter(d)
        writer = self.writers[mode]
        row = {"step": self.step}
        for k, v in losses.items():
            if not k.startswith("_"):
                row[k] = float(v)
                writer.add_scalar(k, row[k], self.step)
        with open(os.path.join(d, "scalars.jsonl"), "a") as f:
            f.write(json.dumps(row) + "\n")
        for j in range(min(4, self.opt.batch_size)):
            for f_id in self.opt.frame_ids:
                if ("color", f_id, 0) in inputs:
                    writer.add_image("color_{}_0/{}".format(f_id, j), inputs[("color", f_id, 0)][j], self.step)
                if f_id != 0 and torch.is_tensor(outputs.get(("color", f_id, 0))):
                    writer.add_image("color_pred_{}_0/{}".format(f_id, j), outputs[("color", f_id, 0)][j], self.step)
            if ("disp", 0) in outputs:
                writer.add_image("disp_mono/{}".format(j), eventlog.colormap(outputs[("disp", 0)][j, 0]), self.step)
            if "depth_mvs" in outputs:
                writer.add_image("disp_mvs/{}".format(j), eventlog.colormap(1 / outputs["depth_mvs"][j].squeeze(0)), self.step)
        writer.flush()

    def save_opts(self):
        """movedepth/trainer.py:796-805."""
        d = os.path.join(self.log_path, "models")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "opt.json"), "w") as f:
            json.dump(dict(self.opt.__dict__), f, indent=2)

    def save_model(self, save_step=False):
        """One <name>.pth state_dict per sub-model + adam.pth (movedepth/trainer.py:807-831); rank 0 only.  The files have
        the reference's formats: plain state_dicts (loadable with strict=True by evaluate_depth.py:115-174) and
        `torch.optim.Adam.state_dict()` for adam.pth (parameter indices in the reference's registration order)."""
        if self.rank != 0:
            return
        tag = "weights_{}".format(self.epoch) if not save_step else "weights_{}_{}".format(self.epoch, self.step)
        if self.epoch == self.opt.num_epochs - 1:
            tag = "last"
        folder = os.path.join(self.log_path, "models", tag)
        os.makedirs(folder, exist_ok=True)
        for name, model in self.models.items():
            sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
            torch.save(sd, os.path.join(folder, "{}.pth".format(name)))
        torch.save(self.adam_state_dict(), os.path.join(folder, "adam.pth"))

    def adam_state_dict(self):
        """The arenas' Adam moments in `torch.optim.Adam.state_dict()` layout (what trainer.py:830-831 saves)."""
        state, groups, idx = {}, [], 0
        step = torch.tensor(float(self.opt_step))
        for a, lr, base in zip(self.arenas, self.current_lrs(), self.base_lrs):
            ids = []
            for i in range(len(a.params)):
                state[idx] = {"step": step.clone(), "exp_avg": a.moment_view(a.exp_avg, i).clone(),
                              "exp_avg_sq": a.moment_view(a.exp_avg_sq, i).clone()}
                ids.append(idx)
                idx += 1
            groups.append({"lr": lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False, "maximize": False,
                           "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                           "decoupled_weight_decay": False, "initial_lr": base, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_adam_state_dict(self, st):
        """Restore the moments and the step count from an Adam state_dict (ours or the reference's)."""
        flat = [(a, j) for a in self.arenas for j in range(len(a.params))]
        n_given = sum(len(g["params"]) for g in st["param_groups"])
        if n_given != len(flat):
            raise ValueError("adam.pth holds %d parameter states, this model has %d" % (n_given, len(flat)))
        step = 0
        for i, (a, j) in enumerate(flat):
            e = st["state"].get(i)
            if e is None:
                continue
            p = a.params[j]
            if tuple(e["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("adam.pth state %d has shape %s, parameter has %s" % (i, tuple(e["exp_avg"].shape), tuple(p.shape)))
            a.moment_view(a.exp_avg, j).copy_(e["exp_avg"])
            a.moment_view(a.exp_avg_sq, j).copy_(e["exp_avg_sq"])
            step = max(step, int(float(e["step"])))
        self.opt_step = step

    def _load_into(self, name, path):
        model = self.models[name]
        have = model.state_dict()
        sd = torch.load(path, map_location="cpu")
        have.update({k: v for k, v in sd.items() if k in have})
        model.load_state_dict(have)              # parameters are views of the arenas: copies land there

    def load_mono_model(self):
        """movedepth/trainer.py:833-844."""
        folder = os.path.expanduser(self.opt.mono_weights_folder)
        for name in ("pose_encoder", "pose", "mono_encoder", "mono_depth"):
            if self.rank == 0:
                print("loading {}".format(name))
            self._load_into(name, os.path.join(folder, "{}.pth".format(name)))

    def load_model(self):
        """movedepth/trainer.py:846-880 (names not in `models`, e.g. the stale defaults, are skipped); the Adam
        moments and step count are restored into the arenas."""
        folder = os.path.expanduser(self.opt.load_weights_folder)
        assert os.path.isdir(folder), "Cannot find folder {}".format(folder)
        if self.rank == 0:
            print("loading model from folder {}".format(folder))
        for name in self.opt.models_to_load:
            path = os.path.join(folder, "{}.pth".format(name))
            if name in self.models and os.path.isfile(path):
                if self.rank == 0:
                    print("Loading {} weights...".format(name))
                self._load_into(name, path)
        adam = os.path.join(folder, "adam.pth")
        if os.path.isfile(adam):
            try:
                if self.rank == 0:
                    print("Loading Adam weights")
                self.load_adam_state_dict(torch.load(adam, map_location=self.device))
            except (ValueError, KeyError, TypeError) as e:
                print("Can't load Adam ({}) - using random".format(e))
        else:
            print("Cannot find Adam weights so Adam is randomly initialized")
