"""Kernel-time breakdown of one steady-state training step (torch.profiler / CUPTI; no replay).
   python tools/profile_step.py [--precision 3xtf32] [--top 40]"""
import argparse
import collections
import os
import re
import sys

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200.options import MonodepthOptions  # noqa: E402
from movedepth_b200.trainer import Trainer, SyntheticKITTI  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="3xtf32")
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--D", type=int, default=96)
    ap.add_argument("--B", type=int, default=6)
    ap.add_argument("--ops", action="store_true", help="also print the per-ATen-op table (self device time): which tensor ops launch the at:: kernels")
    a = ap.parse_args()
    argv = ["--height", "192", "--width", "640", "--num_depth_bins", str(a.D), "--batch_size", str(a.B), "--frame_ids", "0", "-1",
            "--weights_init", "scratch", "--convex_up", "--learning_rate", "2e-4", "--b200_conv_precision", a.precision,
            "--log_dir", "/tmp/mvd_prof"]
    opt = MonodepthOptions().parse(argv)
    torch.manual_seed(0)
    tr = Trainer(opt)
    batch = {k: v.cuda() for k, v in next(iter(SyntheticKITTI(opt, a.B, 1))).items()}
    for _ in range(4):
        tr.train_step(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=a.ops) as prof:
        tr.train_step(batch)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"<.*", "", e.name)[:100]
            agg[name][0] += 1
            agg[name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
            total += e.device_time if hasattr(e, "device_time") else e.cuda_time
    n = sum(c for c, _ in agg.values())
    print("GPU busy %.2f ms over %d kernels/copies in one step (precision=%s)" % (total / 1e3, n, a.precision))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print("%9.1f us %5d %5.1f%%  %s" % (t, c, 100 * t / total, k))
    if a.ops:
        rows = []
        for e in prof.key_averages():
            t = getattr(e, "self_device_time_total", None)
            if t is None:
                t = e.self_cuda_time_total
            if t > 0:
                rows.append((t, e.count, e.key))
        print("\nper-op self device time (top %d)" % a.top)
        for t, c, k in sorted(rows, reverse=True)[:a.top]:
            print("%9.1f us %5d %5.1f%%  %s" % (t, c, 100 * t / total, k[:100]))
        rows = []
        for e in prof.key_averages(group_by_input_shape=True):
            t = getattr(e, "self_device_time_total", None)
            if t is None:
                t = e.self_cuda_time_total
            if t > 0 and e.key.startswith("aten::") and "conv" not in e.key:
                rows.append((t, e.count, e.key, str(e.input_shapes)[:110]))
        crow = []
        for e in prof.key_averages(group_by_input_shape=True):
            t = getattr(e, "device_time_total", None)
            if t is None:
                t = e.cuda_time_total
            if t > 0 and ("convolution" in e.key or e.key in ("_SplitConv", "_PreparedConv", "_SplitConvBackward", "_PreparedConvBackward")):
                crow.append((t, e.count, e.key, str(e.input_shapes)[:130]))
        print("\nconvolution ops by input shape (total device time incl. children, top 70)")
        for t, c, k, sh in sorted(crow, reverse=True)[:70]:
            print("%9.1f us %5d %5.1f%%  %-30s %s" % (t, c, 100 * t / total, k, sh))
        print("\nnon-conv ATen ops by input shape (top 60): the eager tail")
        for t, c, k, sh in sorted(rows, reverse=True)[:60]:
            print("%9.1f us %5d %5.1f%%  %-28s %s" % (t, c, 100 * t / total, k, sh))


if __name__ == "__main__":
    main()
