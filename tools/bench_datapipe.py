"""Throughput of the device data pipeline (movedepth_b200/datapipe.py) at KITTI size: 375x1242 decoded frames -> 192x640
4-scale pyramid + ToTensor (+ colour jitter) for a batch of 6 items x 2 frames.   python tools/bench_datapipe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200.datapipe import DevicePipeline  # noqa: E402


def main():
    dev = "cuda:0"
    B = 6
    pipe = DevicePipeline(192, 640, device=dev)
    K = [[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    g = torch.Generator(device=dev).manual_seed(0)
    frames = {f: torch.randint(0, 256, (B, 375, 1242, 3), dtype=torch.uint8, device=dev, generator=g) for f in (0, -1)}
    for name, kw in (("pyramid + ToTensor", {}), ("+ flip + colour jitter on every item", dict(flip=[1] * B, color_aug=[1] * B))):
        for _ in range(3):
            pipe(frames, K, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            pipe(frames, K, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("%-42s %.2f ms per batch of %d items (2 frames each) = %.0f items/s" % (name, ms, B, B / ms * 1e3))


if __name__ == "__main__":
    main()
