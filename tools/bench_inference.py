"""Batch-1 inference latency of the evaluate_depth path (movedepth/evaluate_depth.py:181-253) at 192x640, D=96:
eager launches vs the CUDA-graph replay of GraphedPredictor.   python tools/bench_inference.py [--arch 18] [--iters 50]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200 import evaluate_depth as ED  # noqa: E402
from movedepth_b200.options import MonodepthOptions  # noqa: E402
from movedepth_b200.trainer import SyntheticKITTI  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", type=int, default=18)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--D", type=int, default=96)
    a = ap.parse_args()
    opt = MonodepthOptions().parse(["--height", "192", "--width", "640", "--num_depth_bins", str(a.D), "--batch_size", "1",
                                    "--res_arch", str(a.arch), "--weights_init", "scratch", "--convex_up", "--frame_ids", "0", "-1"])
    torch.manual_seed(0)
    pred = ED.DepthPredictor(opt)
    batch = {k: v.cuda() for k, v in next(iter(SyntheticKITTI(opt, 1, 1, smooth=True))).items()}
    for _ in range(5):
        pred.predict(batch)
    gp = ED.GraphedPredictor(pred, batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for name, fn in (("eager", pred.predict), ("cuda_graph", gp.predict)):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            fn(batch)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.iters
    print("batch-1 inference latency, ResNet%d 2-frame 192x640 D=%d (3xTF32 convolutions): eager %.2f ms, CUDA graph %.2f ms (%.0f frames/s)"
          % (a.arch, a.D, res["eager"], res["cuda_graph"], 1e3 / res["cuda_graph"]))


if __name__ == "__main__":
    main()
