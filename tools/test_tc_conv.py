"""Bring-up check of the tcgen05 16->16 conv: numerics vs torch conv3d (fp64, CPU) for both descriptor variants,
then timing at the BASELINE config-2 volume.   python tools/test_tc_conv.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200 import ops  # noqa: E402


def main():
    dev = "cuda:0"
    gen = torch.Generator().manual_seed(4)
    for shape in (() if "--wgrad-only" in sys.argv else ((1, 6, 7, 32), (1, 5, 11, 45), (2, 26, 24, 80))):
        B, D, H, W = shape
        x = torch.randn(B, 16, D, H, W, generator=gen)
        w = torch.randn(16, 16, 3, 3, 3, generator=gen) * 0.1
        yo = F.conv3d(x.double(), w.double(), padding=1).float()
        go = torch.nn.grad.conv3d_input(x.shape, w.double(), x.double(), padding=1).float()   # dgrad with gy := x
        xg = x.to(dev).contiguous(memory_format=torch.channels_last_3d)
        wg = w.to(dev)
        for flags in (0,):
            for passes in (1, 3):
                y = ops.c16c16_conv_tc(xg, wg, 0, passes, flags)
                torch.cuda.synchronize()
                e = float((y.cpu() - yo).abs().max() / yo.abs().max())
                g = ops.c16c16_conv_tc(xg, wg, 1, passes, flags)
                torch.cuda.synchronize()
                eg = float((g.cpu() - go).abs().max() / go.abs().max())
                print("shape %s flags %d passes %d: fwd rel err %.2e, dgrad rel err %.2e" % (shape, flags, passes, e, eg), flush=True)
    # weight gradient
    for shape in ((1, 6, 7, 32), (1, 5, 11, 46), (2, 26, 24, 80), (1, 9, 20, 160)):
        B, D, H, W = shape
        x = torch.randn(B, 16, D, H, W, generator=gen)
        gy = torch.randn(B, 16, D, H, W, generator=gen)
        gwo = torch.nn.grad.conv3d_weight(x.double(), (16, 16, 3, 3, 3), gy.double(), padding=1).float()
        xg = x.to(dev).contiguous(memory_format=torch.channels_last_3d)
        gg = gy.to(dev).contiguous(memory_format=torch.channels_last_3d)
        gw = ops.c16c16_wgrad_tc(gg, xg)
        torch.cuda.synchronize()
        gwf = ops.c16c16_wgrad(gg, xg)
        print("wgrad shape %s: tc rel err %.2e, ffma2 rel err %.2e" % (shape, float((gw.cpu() - gwo).abs().max() / gwo.abs().max()),
                                                                       float((gwf.cpu() - gwo).abs().max() / gwo.abs().max())), flush=True)
    # timing
    x = torch.randn(6, 16, 96, 48, 160, device=dev).contiguous(memory_format=torch.channels_last_3d)
    w = torch.randn(16, 16, 3, 3, 3, device=dev) * 0.1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cases = [("wgrad tc", lambda: ops.c16c16_wgrad_tc(x, x)), ("wgrad ffma2", lambda: ops.c16c16_wgrad(x, x)),
             ("tc 3-pass", lambda: ops.c16c16_conv_tc(x, w, 0, 3)), ("tc 1-pass", lambda: ops.c16c16_conv_tc(x, w, 1, 1)),
             ("mma.sync 1-pass", lambda: ops.c16c16_conv(x, w, 1, 1))]
    for dbg, nm in ((2, 'no MMA'), (4, 'no stores'), (8, 'no split'), (14, 'no MMA/stores/split')) if '--breakdown' in sys.argv else ():
        cases.append(("tc 3-pass dbg %s" % nm, lambda dbg=dbg: ops.c16c16_conv_tc(x, w, 0, 3, dbg)))
        cases.append(("tc 1-pass dbg %s" % nm, lambda dbg=dbg: ops.c16c16_conv_tc(x, w, 1, 1, dbg)))
    for name, fn in cases:
        ts = []
        for i in range(8):
            flush.fill_(i & 1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print("%s: median %.1f us" % (name, ts[len(ts) // 2] * 1e3), flush=True)


if __name__ == "__main__":
    main()
