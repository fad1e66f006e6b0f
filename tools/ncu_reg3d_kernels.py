"""One reg3d forward+backward at BASELINE config 2 for `ncu -k regex:...` captures of the own kernels (no torch profiler).
   ncu --set full --clock-control none -k regex:"c16o1|c16c16|bn_" -s <skip> -c <count> python tools/ncu_reg3d_kernels.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200 import networks as PN  # noqa: E402
from movedepth_b200 import precision as PR  # noqa: E402

PR.set_policy("3xtf32")
torch.manual_seed(0)
reg = PN.reg3d(16, 16, 3).to("cuda:0")
vol = torch.randn(6, 16, 96, 48, 160, device="cuda:0").contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    reg.forward_volume(vol).square().sum().backward()
torch.cuda.synchronize()
