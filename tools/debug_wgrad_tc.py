"""Tap-by-tap check of the tcgen05 weight gradient against torch's fp64 conv3d_weight (which reference tap, transposed or
not, does each computed [co x ci] block match best?) -- the bring-up tool for the MN-major descriptor layout.
   python tools/debug_wgrad_tc.py"""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200 import ops, _lib
dev = "cuda:0"
gen = torch.Generator().manual_seed(4)
B, D, H, W = 1, 6, 7, 32
x = torch.randn(B, 16, D, H, W, generator=gen)
gy = torch.randn(B, 16, D, H, W, generator=gen)
ref = torch.nn.grad.conv3d_weight(x.double(), (16, 16, 3, 3, 3), gy.double(), padding=1).float()   # [co,ci,kd,kh,kw]
xg = x.to(dev).contiguous(memory_format=torch.channels_last_3d); gg = gy.to(dev).contiguous(memory_format=torch.channels_last_3d)
got = ops.c16c16_wgrad_tc(gg, xg)
torch.cuda.synchronize()
got = got.cpu()
print("norm ref %.3f got %.3f  nan %d  rel err %.3e" % (ref.norm(), got.norm(), int(torch.isnan(got).sum()), float((got - ref).abs().max() / ref.abs().max())))
r = ref.reshape(16, 16, 27); g = got.reshape(16, 16, 27)
for t in range(27):
    a = g[:, :, t]
    best = None
    for t2 in range(27):
        for name, b in (("same", r[:, :, t2]), ("T", r[:, :, t2].t())):
            e = float((a - b).norm() / (b.norm() + 1e-9))
            if best is None or e < best[0]:
                best = (e, t2, name)
    print("tap %2d (kd %d kh %d kw %d): |got| %.2f |ref| %.2f best match ref tap %2d %s err %.3f" % (t, t // 9, (t // 3) % 3, t % 3, float(a.norm()), float(r[:, :, t].norm()), best[1], best[2], best[0]))
