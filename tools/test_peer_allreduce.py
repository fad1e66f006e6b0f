"""2+ GPU check of the peer-memory all-reduce (csrc/peer.cu) against NCCL, eager and inside a CUDA graph.
   torchrun --nproc-per-node 2 tools/test_peer_allreduce.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200 import norm as NM  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    px = NM.PeerExchange.get(dev)
    assert px is not None, "peer exchange unavailable"
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    worst = 0.0
    for it in range(300):
        n = [16, 64, 256, 1024, 2048][it % 5]
        v = torch.randn(n, device=dev, dtype=torch.float64, generator=g)
        want = v.clone()
        dist.all_reduce(want)
        px.allreduce_(v)
        worst = max(worst, float((v - want).abs().max()))
    torch.cuda.synchronize()
    # inside a CUDA graph
    bufs = [torch.randn(128, device=dev, dtype=torch.float64, generator=g) for _ in range(50)]
    want = []
    for b in bufs:
        w = b.clone()
        dist.all_reduce(w)
        want.append(w)
    s = torch.cuda.Stream()
    work = [b.clone() for b in bufs]
    s.wait_stream(torch.cuda.current_stream())          # after the clones were enqueued
    with torch.cuda.stream(s):
        for w in work:
            px.allreduce_(w)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    serr = max(float((a - b).abs().max()) for a, b in zip(work, want))
    gerr = 0.0
    graph = torch.cuda.CUDAGraph()
    static = [b.clone() for b in bufs]
    outs = [torch.empty_like(b) for b in bufs]
    with torch.cuda.graph(graph, stream=s):
        for a, o in zip(static, outs):
            o.copy_(a)
            px.allreduce_(o)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    gerr = max(gerr, max(float((a - b).abs().max()) for a, b in zip(outs, want)))
    # latency
    v = torch.randn(128, device=dev, dtype=torch.float64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(200):
        px.allreduce_(v)
    e1.record()
    torch.cuda.synchronize()
    t_peer = e0.elapsed_time(e1) / 200 * 1e3
    e0.record()
    for _ in range(200):
        dist.all_reduce(v)
    e1.record()
    torch.cuda.synchronize()
    t_nccl = e0.elapsed_time(e1) / 200 * 1e3
    print("rank %d/%d: eager max err %.2e, side-stream max err %.2e, graph max err %.2e, peer %.1f us/call, nccl %.1f us/call (eager launches)" % (
        rank, world, worst, serr, gerr, t_peer, t_nccl), flush=True)
    # nn.SyncBatchNorm through bn_act on real peers: each rank holds a slice of one global batch; outputs / gradients must
    # equal BatchNorm over the whole batch (fp64), for the one-kernel form (exchange at the grid barrier) and the three-kernel
    # form (exchange by the last block of the reductions)
    import torch.nn as nn
    gen = torch.Generator().manual_seed(5)
    for shape, one_kernel in [((4 * world, 64, 24, 80), True), ((4 * world, 64, 24, 80), False), ((2 * world, 16, 6, 24, 80), True),
                              ((2 * world, 512, 6, 20), True)]:
        NM.fuse_bytes = (1 << 30) if one_kernel else 0
        C, nb = shape[1], shape[0] // world
        x = torch.randn(shape, generator=gen) * 1.5 + 0.3
        gy = torch.randn(shape, generator=gen)
        ref = (nn.BatchNorm3d if len(shape) == 5 else nn.BatchNorm2d)(C).double()
        xo = x.double().requires_grad_(True)
        yo = torch.relu(ref(xo))
        (yo * gy.double()).sum().backward()
        bn = nn.SyncBatchNorm(C).to(dev)
        fmt = torch.channels_last_3d if len(shape) == 5 else torch.channels_last
        xs = x[rank * nb:(rank + 1) * nb].to(dev).contiguous(memory_format=fmt).requires_grad_(True)
        for rep in range(3):                                          # repeated: the shared workspace must come back clean
            bn.zero_grad()
            xs.grad = None
            y = NM.bn_act(bn, xs, relu=True)
            (y * gy[rank * nb:(rank + 1) * nb].to(dev)).sum().backward()
        sl = slice(rank * nb, (rank + 1) * nb)
        ey = float((y.detach().cpu() - yo[sl].detach().float()).abs().max())
        eg = float((xs.grad.cpu() - xo.grad[sl].float()).abs().max()) / float(xo.grad.abs().max())
        gw = bn.weight.grad.clone()
        dist.all_reduce(gw)
        ew = float((gw.cpu() - ref.weight.grad.float()).abs().max()) / float(ref.weight.grad.abs().max())
        assert ey < 5e-5 and eg < 2e-4 and ew < 2e-4, (shape, one_kernel, ey, eg, ew)
        NM.check_workspaces()
        if rank == 0:
            print("SyncBatchNorm %s %s: y err %.1e, gx err %.1e (rel), sum of rank gw err %.1e (rel)" % (
                shape, "one-kernel" if one_kernel else "three-kernel", ey, eg, ew), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
