"""Stand-alone driver for the cost-volume kernels at BASELINE config-2 size (for ncu / timing).

    python tools/profile_costvol.py [--iters 20] [--pose forward|sideways|stress] [--bwd] [--layout 0|1]

Prints the CUDA-event time per launch with an L2 flush between launches."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as C  # noqa: E402
from movedepth_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--pose", default="forward")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--layout", type=int, default=0)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--B", type=int, default=6)
    ap.add_argument("--h", type=int, default=48)
    ap.add_argument("--w", type=int, default=160)
    ap.add_argument("--D", type=int, default=96)
    ap.add_argument("--tscale", type=float, default=1.0, help="scale the pose translation (longer epipolar footprints)")
    a = ap.parse_args()
    c = C.case_costvol(a.pose, B=a.B, h=a.h, w=a.w, D=a.D)
    if a.tscale != 1.0:
        c["pose"] = c["pose"].clone()
        c["pose"][:, 0, :3, 3] *= a.tscale
    dev = "cuda:0"
    ref = c["ref"].to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(a.bwd)
    src = c["src"].to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(a.bwd)
    geo = [c[k].to(dev) for k in ("K", "invK")] + [c["pose"][:, 0].to(dev)]
    prior, ratio = c["prior"].to(dev), c["ratio"].to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nbytes = 4 * a.B * a.h * a.w * (64 + 1 + a.D * 16)
    tf, tb = [], []
    for i in range(a.iters + 3):
        flush.fill_(i & 1)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        out = ops.costvol_grouped(ref, src, *geo, prior=prior, ratio=ratio, layout=a.layout, flags=a.flags)
        e[1].record()
        if a.bwd:
            gout = torch.ones_like(out)
            flush.fill_(i & 1)
            e[2].record()
            out.backward(gout)
            e[3].record()
            ref.grad = src.grad = None
        torch.cuda.synchronize()
        if i >= 3:
            tf.append(e[0].elapsed_time(e[1]))
            if a.bwd:
                tb.append(e[2].elapsed_time(e[3]))
    tf.sort()
    med = tf[len(tf) // 2]
    print("costvol fwd  pose=%s layout=%d flags=%d: median %.1f us, min %.1f us -> %.0f GB/s algorithmic (%.1f MB)"
          % (a.pose, a.layout, a.flags, med * 1e3, tf[0] * 1e3, nbytes / med / 1e6, nbytes / 1e6))
    if tb:
        tb.sort()
        print("costvol bwd (memsets + kernel): median %.1f us, min %.1f us" % (tb[len(tb) // 2] * 1e3, tb[0] * 1e3))


if __name__ == "__main__":
    main()
