"""Which aten ops own the GPU time of one eager training step?  (torch.profiler, grouped by operator and by
python call site of movedepth_b200)   python tools/profile_ops.py [--top 40]"""
import argparse
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200.options import MonodepthOptions  # noqa: E402
from movedepth_b200.trainer import Trainer, SyntheticKITTI  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    argv = ["--height", "192", "--width", "640", "--num_depth_bins", "96", "--batch_size", "6", "--frame_ids", "0", "-1",
            "--weights_init", "scratch", "--convex_up", "--learning_rate", "2e-4", "--log_dir", "/tmp/mvd_prof"]
    opt = MonodepthOptions().parse(argv)
    torch.manual_seed(0)
    tr = Trainer(opt)
    batch = {k: v.cuda() for k, v in next(iter(SyntheticKITTI(opt, 6, 1))).items()}
    for _ in range(4):
        tr.train_step(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        tr.train_step(batch)
        torch.cuda.synchronize()
    ka = prof.key_averages()
    rows = sorted(ka, key=lambda e: -e.self_device_time_total)
    total = sum(e.self_device_time_total for e in rows)
    print("self GPU time by operator: total %.2f ms" % (total / 1e3))
    for e in rows[:a.top]:
        print("%9.1f us %5d %5.1f%%  %s" % (e.self_device_time_total, e.count, 100 * e.self_device_time_total / total, e.key[:90]))
    print("\nby call site (innermost movedepth_b200 frame):")
    ks = prof.key_averages(group_by_stack_n=12)
    site = {}
    for e in ks:
        if e.self_device_time_total <= 0:
            continue
        fr = [s for s in e.stack if "movedepth_b200" in s]
        key = (fr[0].split("movedepth_b200/")[-1] if fr else "<other>")[:80]
        site[key] = site.get(key, 0) + e.self_device_time_total
    for k, v in sorted(site.items(), key=lambda kv: -kv[1])[:a.top]:
        print("%9.1f us %5.1f%%  %s" % (v, 100 * v / total, k))


if __name__ == "__main__":
    main()
