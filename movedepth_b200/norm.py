"""Training-mode BatchNorm (+ReLU, + residual add) on the hand-written channels-last kernels of csrc/bn.cu.

`bn_act(bn, x, relu=False, residual=None)` is what the conv blocks call instead of `relu(bn(x) [+ residual])`.
The BatchNorm modules themselves stay stock `nn.BatchNorm{2,3}d` / `nn.SyncBatchNorm` objects (same parameters,
buffers and state-dict keys as the reference: movedepth/networks/resnet_encoder.py:175-182, 453-475; torchvision
ResNet blocks).  Under `nn.SyncBatchNorm` (data-parallel training, movedepth/trainer.py:69-129) the per-channel
fp64 sums are all-reduced between the statistics and the apply kernels -- one small NCCL call per direction instead of
SyncBatchNorm's all_gather + gather_stats pipeline.  Evaluation mode and CPU tensors use torch's own batch_norm.
"""
import ctypes
import os
import types

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

enabled = True            # set False to route everything through torch's batch_norm (A/B measurements)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _fmt(t):
    return torch.channels_last_3d if t.dim() == 5 else torch.channels_last


class PeerExchange:
    """All-reduce of the small fp64 statistic vectors by our own kernel over NVLink peer memory
    (csrc/peer.cu) instead of NCCL.  One symmetric buffer per process (torch symmetric memory: CUDA VMM
    allocations mapped into every rank of the node)."""
    NMAX = 2048
    _inst = {}               # one exchange buffer (and epoch counter) per channel: concurrent streams must not share one
    _failed = False

    def __init__(self, device):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        nbytes = _lib.lib().mvd_peer_allreduce_buffer_bytes(self.world, self.NMAX)
        self.buf = symm.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize()
        dist.barrier()

    @classmethod
    def get(cls, device, chan=0):
        if chan not in cls._inst and not cls._failed:
            try:
                cls._inst[chan] = cls(device)
            except Exception as e:                      # no peer access / symmetric memory: NCCL does the exchange
                cls._failed = True
                print("movedepth_b200: peer-memory SyncBN exchange unavailable (%s: %s); using NCCL" % (type(e).__name__, e))
        return cls._inst.get(chan)

    def allreduce_(self, v):
        rc = _lib.lib().mvd_peer_allreduce_f64(_p(v), _p(v), v.numel(), _p(self.ptrs), self.rank, self.world, self.NMAX, _stream())
        _lib.check(rc, "mvd_peer_allreduce_f64")
        _count_launch(1)


peer_exchange = True      # SyncBN statistics over NVLink peer memory when available (else NCCL all_reduce)
channel = 0               # exchange channel of the BatchNorm layers being built right now: the trainer runs its two independent
                          # autograd graphs on two streams, each with its own exchange buffer (the epochs of one buffer must
                          # advance in the same order on every rank)


def _peer(device, C, chan=0):
    """(peers table pointer, rank, world, nmax) for the in-kernel exchange, or None -> NCCL all_reduce after the kernel."""
    px = PeerExchange.get(device, chan) if (peer_exchange and 2 * C <= PeerExchange.NMAX) else None
    return px


fuse_bytes = int(float(os.environ.get("MVD_BN_FUSE_MB", "64")) * (1 << 20))
"""Activations up to this size take the one-kernel forward / backward (csrc/bn.cu: reduce, grid barrier with the SyncBN exchange,
apply from the L2); larger ones stream through the separate reduction / apply kernels at full-grid bandwidth."""
_workspaces = {}


def workspace(device, chan=0):
    """The zero-initialised statistics workspace of the one-kernel BatchNorm (handed back zeroed by every call).  One per
    exchange channel: layers on one stream share it, the trainer's two concurrent streams use channels 0 and 1.  Created
    outside any CUDA-graph capture (`Trainer.__init__` touches both channels)."""
    key = (torch.device(device).index or 0, chan)
    ws = _workspaces.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("movedepth_b200.norm.workspace(%r, %d) must be created before CUDA-graph capture" % (device, chan))
        ws = _workspaces[key] = torch.zeros(_lib.lib().mvd_bn_workspace_doubles(), device=device, dtype=torch.float64)
    return ws


def check_workspaces():
    """Raises if a grid barrier of the one-kernel BatchNorm ever timed out (control word 3 is sticky).  Synchronises: the
    trainer calls it on its log steps."""
    for (dev, chan), ws in _workspaces.items():
        if int(ws[-2:].view(torch.int32)[3]) != 0:
            raise RuntimeError("movedepth_b200: a BatchNorm grid barrier timed out on cuda:%d (channel %d)" % (dev, chan))


def _count_launch(n):
    from . import ops
    ops.launch_counter["n"] += n


class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, momentum, eps, relu, sync, post=False, nbt=None,
                sums=None):
        L = _lib.lib()
        C = x.shape[1]
        xc = x.contiguous(memory_format=_fmt(x))
        M = xc.numel() // C
        rc_ = residual.contiguous(memory_format=_fmt(x)) if residual is not None else None
        ctx.channel = channel
        px = _peer(x.device, C, channel) if sync else None
        precomputed = sums is not None and sums.numel() >= 2 * C       # the producing conv's epilogue already summed its output
        post = bool(post and residual is not None)         # post: relu(bn(x)) + residual (U-Net skip); else relu(bn(x) + residual)
        flags = int(relu) | (2 if post else 0)
        count = float(M) * (dist.get_world_size() if sync else 1)
        stats = torch.empty(4 * C, device=x.device, dtype=torch.float32)
        y = torch.empty_like(xc)
        fused = (not precomputed) and xc.numel() * 4 <= fuse_bytes and (px is not None or not sync)
        if fused:                                          # one launch: reduce, barrier (+ exchange), finalize, apply
            _lib.check(L.mvd_bn_fwd_fused(_p(xc), _p(rc_), _p(weight), _p(bias), _p(running_mean), _p(running_var), _p(nbt),
                                          float(momentum), float(eps), count, _p(stats), _p(y), M, C, flags,
                                          _p(workspace(x.device, channel)), _p(px.ptrs) if px else _p(None), px.rank if px else 0,
                                          px.world if px else 1, px.NMAX if px else 0, _stream()), "mvd_bn_fwd_fused")
            _count_launch(1)
        else:
            if not precomputed:
                sums = torch.empty(2 * C + 1, device=x.device, dtype=torch.float64)      # [sum x, sum x^2, arrival counter]
                _lib.check(L.mvd_bn_stats(_p(xc), M, C, _p(sums), _p(px.ptrs) if px else _p(None), px.rank if px else 0,
                                          px.world if px else 1, px.NMAX if px else 0, _stream()), "mvd_bn_stats")
            if sync:
                if precomputed and px is not None:            # fused statistics: the exchange is its own (single-CTA) kernel
                    px.allreduce_(sums[:2 * C])
                elif px is None:                              # no peer memory: NCCL exchanges the sums
                    dist.all_reduce(sums[:2 * C])
            _lib.check(L.mvd_bn_finalize(_p(sums), count, _p(weight), _p(bias), _p(running_mean), _p(running_var), float(momentum),
                                         float(eps), _p(stats), C, _p(nbt), _stream()), "mvd_bn_finalize")
            _lib.check(L.mvd_bn_apply(_p(xc), _p(rc_), _p(stats), _p(y), M, C, flags, _stream()), "mvd_bn_apply")
            _count_launch(4)
        # the ReLU mask is the sign of bn(x), recomputed from x in the backward, unless a residual was added BEFORE the ReLU
        ctx.save_for_backward(xc, y if (relu and residual is not None and not post) else None, stats, weight)
        ctx.cfg = (M, C, count, bool(relu), bool(sync), residual is not None, post, fused)
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, y, stats, weight = ctx.saved_tensors
        M, C, count, relu, sync, has_res, post, fused = ctx.cfg
        L = _lib.lib()
        gy = gy.contiguous(memory_format=_fmt(xc))
        px = _peer(xc.device, C, ctx.channel) if sync else None
        local2 = torch.empty(2 * C, device=xc.device, dtype=torch.float64) if px else None
        gx = torch.empty_like(xc)
        gres = torch.empty_like(xc) if (has_res and not post and ctx.needs_input_grad[3]) else None
        gw = gb = None
        if not sync:
            gw = torch.empty(C, device=xc.device, dtype=torch.float32)
            gb = torch.empty(C, device=xc.device, dtype=torch.float32)
        if fused:
            _lib.check(L.mvd_bn_bwd_fused(_p(gy), _p(xc), _p(y), _p(stats), _p(weight), count, _p(gx), _p(gres), _p(gw), _p(gb),
                                          _p(local2), M, C, int(relu), _p(workspace(xc.device, ctx.channel)),
                                          _p(px.ptrs) if px else _p(None), px.rank if px else 0, px.world if px else 1,
                                          px.NMAX if px else 0, _stream()), "mvd_bn_bwd_fused")
            _count_launch(1)
            if sync:                              # parameter gradients are this rank's own sums (DDP averages them later)
                gb, gw = local2[:C].float(), local2[C:].float()
        else:
            sums2 = torch.empty(2 * C + 1, device=xc.device, dtype=torch.float64)
            _lib.check(L.mvd_bn_bwd_reduce(_p(gy), _p(xc), _p(y), _p(stats), _p(sums2), _p(local2), M, C, int(relu),
                                           _p(px.ptrs) if px else _p(None), px.rank if px else 0, px.world if px else 1,
                                           px.NMAX if px else 0, _stream()), "mvd_bn_bwd_reduce")
            if sync:
                if px is None:
                    gb, gw = sums2[:C].float(), sums2[C:2 * C].float()
                    dist.all_reduce(sums2[:2 * C])
                else:                             # the reduction kernel exchanged in place and kept the local sums aside
                    gb, gw = local2[:C].float(), local2[C:].float()
            _lib.check(L.mvd_bn_bwd_apply(_p(gy), _p(xc), _p(y), _p(stats), _p(weight), _p(sums2), count, _p(gx), _p(gres),
                                          _p(None if sync else gw), _p(None if sync else gb), M, C, int(relu), _stream()),
                       "mvd_bn_bwd_apply")
            _count_launch(3)
        if post and ctx.needs_input_grad[3]:
            gres = gy                                 # the skip was added after the ReLU: its gradient is gy itself
        if weight is None or not ctx.needs_input_grad[1]:
            gw = None
        if not ctx.needs_input_grad[2]:
            gb = None
        return gx, gw, gb, gres, None, None, None, None, None, None, None, None, None


def _fusable(bn, x):
    C = x.shape[1]
    return (enabled and bn.training and x.is_cuda and x.dtype == torch.float32 and x.dim() in (4, 5) and bn.track_running_stats
            and bn.momentum is not None and 4 <= C <= 1024 and (C & (C - 1)) == 0 and x.numel() > 0)


def bn_act(bn, x, relu=False, residual=None, post=False, sums=None):
    """relu(bn(x) + residual), or relu(bn(x)) + residual with post=True, with `bn` a BatchNorm2d / BatchNorm3d /
    SyncBatchNorm module."""
    if not _fusable(bn, x):
        y = bn(x)
        if post:
            y = F.relu(y) if relu else y
            return y + residual if residual is not None else y
        if residual is not None:
            y = y + residual
        return F.relu(y, inplace=True) if relu else y
    sync = isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    # num_batches_tracked is bumped inside the finalize kernel (95 one-element add kernels per step otherwise)
    return _BNAct.apply(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu, sync, post,
                        bn.num_batches_tracked, sums)


# ---- torchvision ResNet blocks: same modules / parameters, forward routed through bn_act -----------------------------
def _basic_block_forward(self, x):
    identity = x if self.downsample is None else _downsample(self.downsample, x)
    out = bn_act(self.bn1, self.conv1(x), relu=True)
    return bn_act(self.bn2, self.conv2(out), relu=True, residual=identity)


def _bottleneck_forward(self, x):
    identity = x if self.downsample is None else _downsample(self.downsample, x)
    out = bn_act(self.bn1, self.conv1(x), relu=True)
    out = bn_act(self.bn2, self.conv2(out), relu=True)
    return bn_act(self.bn3, self.conv3(out), relu=True, residual=identity)


def _downsample(seq, x):
    if len(seq) == 2 and isinstance(seq[1], nn.modules.batchnorm._BatchNorm):
        return bn_act(seq[1], seq[0](x))
    return seq(x)


def adopt(module):
    """Route the residual blocks of a torchvision ResNet through bn_act (in place; parameters untouched)."""
    from torchvision.models.resnet import BasicBlock, Bottleneck
    for m in module.modules():
        if isinstance(m, BasicBlock):
            m.forward = types.MethodType(_basic_block_forward, m)
        elif isinstance(m, Bottleneck):
            m.forward = types.MethodType(_bottleneck_forward, m)
    return module
