"""Micro-benchmark: how fast / how accurate are the cuDNN conv stacks of the hot path on this GPU
under different arithmetic policies (fp32 SIMT, TF32, 3xTF32 split emulation, channels-last)?

    python tools/bench_convs.py [--B 6] [--D 96]
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from movedepth_b200 import networks as PN  # noqa: E402


def tf32_round(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def split3(x, dim):
    hi = tf32_round(x)
    return torch.cat([hi, x - hi, hi], dim)


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def conv_case(name, x, w, stride, transposed, iters=5):
    """single conv layer fwd+bwd under each policy, plus accuracy vs fp64"""
    f = (lambda a, b: F.conv_transpose3d(a, b, stride=stride, padding=1, output_padding=stride - 1)) if transposed else \
        (lambda a, b: F.conv3d(a, b, stride=stride, padding=1))
    ref = f(x.double(), w.double())
    res = {}
    for pol in ("fp32", "tf32", "3xtf32", "tf32_cl", "3xtf32_cl"):
        torch.backends.cudnn.allow_tf32 = pol != "fp32"
        cl = pol.endswith("_cl")
        xx = x.contiguous(memory_format=torch.channels_last_3d) if cl else x
        ww = w.contiguous(memory_format=torch.channels_last_3d) if cl else w
        xx = xx.clone().requires_grad_(True)
        ww = ww.clone().requires_grad_(True)
        if pol.startswith("3xtf32"):
            wdim = 0 if transposed else 1

            def run():
                y = f(split3(xx, 1), torch.cat([tf32_round(ww), tf32_round(ww), ww - tf32_round(ww)], wdim))
                y.sum().backward()
                return y
        else:
            def run():
                y = f(xx, ww)
                y.sum().backward()
                return y
        y = run()
        err = float((y.double() - ref).abs().max() / ref.abs().max())
        res[pol] = (timeit(run, iters), err)
    print("%-28s " % name + "  ".join("%s %.2f ms (err %.1e)" % (k, v[0], v[1]) for k, v in res.items()), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=6)
    ap.add_argument("--D", type=int, default=96)
    a = ap.parse_args()
    dev = "cuda:0"
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    B, D, h, w = a.B, a.D, 48, 160
    g = torch.Generator(device=dev).manual_seed(0)

    def rn(*s):
        return torch.randn(*s, device=dev, generator=g)
    conv_case("conv0 16->16 s1 full", rn(B, 16, D, h, w), rn(16, 16, 3, 3, 3) * 0.05, 1, False)
    conv_case("conv1 16->32 s2", rn(B, 16, D, h, w), rn(32, 16, 3, 3, 3) * 0.05, 2, False)
    conv_case("conv2 32->32 s1 half", rn(B, 32, D // 2, h // 2, w // 2), rn(32, 32, 3, 3, 3) * 0.05, 1, False)
    conv_case("conv4 64->64 s1 quarter", rn(B, 64, D // 4, h // 4, w // 4), rn(64, 64, 3, 3, 3) * 0.05, 1, False)
    conv_case("conv6 128->128 s1 eighth", rn(B, 128, D // 8, h // 8, w // 8), rn(128, 128, 3, 3, 3) * 0.05, 1, False)
    conv_case("conv11T 32->16 s2", rn(B, 32, D // 2, h // 2, w // 2), rn(32, 16, 3, 3, 3) * 0.05, 2, True)
    conv_case("prob 16->1 s1 full", rn(B, 16, D, h, w), rn(1, 16, 3, 3, 3) * 0.05, 1, False)

    # whole sub-networks, fwd+bwd
    def net_time(name, net, inp, fmt=None):
        out = {}
        for pol in ("fp32", "tf32"):
            torch.backends.cudnn.allow_tf32 = pol == "tf32"
            x = inp.clone()
            if fmt is not None:
                x = x.contiguous(memory_format=fmt)

            def run():
                y = net(x)
                y = y[-1] if isinstance(y, (list, tuple)) else y
                if isinstance(y, dict):
                    y = sum(v.sum() for v in y.values())
                y.sum().backward()
            out[pol] = timeit(run, 3)
        print("%-28s " % name + "  ".join("%s %.2f ms" % kv for kv in out.items()), flush=True)

    reg = PN.reg3d(16, 16, 3).to(dev)
    vol = rn(B, 16, D, h, w)
    net_time("reg3d NCDHW", lambda v: reg.forward_volume(v), vol)
    reg_cl = PN.reg3d(16, 16, 3).to(dev).to(memory_format=torch.channels_last_3d)
    net_time("reg3d NDHWC", lambda v: reg_cl.forward_volume(v), vol, torch.channels_last_3d)
    enc = PN.ResnetEncoder(18, False).to(dev)
    img = torch.rand(B, 3, 192, 640, device=dev)
    net_time("resnet18 enc NCHW", enc, img)
    enc_cl = PN.ResnetEncoder(18, False).to(dev).to(memory_format=torch.channels_last)
    net_time("resnet18 enc NHWC", enc_cl, img, torch.channels_last)
    dec = PN.DepthDecoder(enc.num_ch_enc).to(dev)
    net_time("enc+decoder NCHW", lambda x: dec(enc(x)), img)
    dec_cl = PN.DepthDecoder(enc.num_ch_enc).to(dev).to(memory_format=torch.channels_last)
    net_time("enc+decoder NHWC", lambda x: dec_cl(enc_cl(x)), img, torch.channels_last)
    fpn = PN.FPN4(8, 2).to(dev)
    net_time("FPN4 NCHW", lambda x: fpn(x)[0], img)
    fpn_cl = PN.FPN4(8, 2).to(dev).to(memory_format=torch.channels_last)
    net_time("FPN4 NHWC", lambda x: fpn_cl(x)[0], img, torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        net_time("reg3d NDHWC bf16-autocast", lambda v: reg_cl.forward_volume(v), vol, torch.channels_last_3d)


if __name__ == "__main__":
    main()
