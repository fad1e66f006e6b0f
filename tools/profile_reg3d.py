"""Kernel-by-kernel timeline of one reg3d forward+backward at BASELINE config 2 (volume [6,16,96,48,160],
channels-last-3d, default precision policy), in launch order.   python tools/profile_reg3d.py"""
import os
import re
import sys

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200 import networks as PN  # noqa: E402
from movedepth_b200 import precision as PR  # noqa: E402


def main():
    pol = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
    dev = "cuda:0"
    torch.backends.cudnn.benchmark = True
    PR.set_policy(pol)
    torch.manual_seed(0)
    reg = PN.reg3d(16, 16, 3).to(dev)
    vol = torch.randn(6, 16, 96, 48, 160, device=dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    for _ in range(3):
        reg.forward_volume(vol).sum().backward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        reg.forward_volume(vol).square().sum().backward()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    total = sum(e.device_time for e in evs)
    print("reg3d fwd+bwd (%s): %.2f ms GPU busy, %d kernels" % (pol, total / 1e3, len(evs)))
    for e in evs:
        print("%9.1f us  %s" % (e.device_time, re.sub(r"\(.*", "", e.name)[:110]))


if __name__ == "__main__":
    main()
