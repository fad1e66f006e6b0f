"""End-to-end parity table: Trainer (GPU) vs the reference's golden step outputs, per precision policy.
Prints, per step case, the fraction of pixels within 1e-3 relative for the depth maps and the max
relative error of the mono disparity.   python tools/parity_report.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as C  # noqa: E402
from _weights import fill_deterministic  # noqa: E402
from movedepth_b200.options import MonodepthOptions  # noqa: E402
from movedepth_b200.trainer import Trainer  # noqa: E402


def main():
    gold_dir = os.path.join(ROOT, "tests", "golden")
    print("%-10s %-11s %9s %9s %9s %11s %11s" % ("case", "policy", "depth_mvs", "masked", "fused", "disp0 maxrel", "loss rel"))
    for name, cfg in C.STEP_CASES.items():
        gold = dict(np.load(os.path.join(gold_dir, "step_%s.npz" % name)))
        for pol in ("fp32", "3xtf32", "tf32"):
            argv = ["--height", str(cfg["H"]), "--width", str(cfg["W"]), "--num_depth_bins", str(cfg["D"]), "--batch_size",
                    str(cfg["B"]), "--res_arch", str(cfg.get("arch", 18)), "--weights_init", "scratch", "--convex_up",
                    "--b200_conv_precision", pol, "--log_dir", "/tmp/mvd_parity", "--frame_ids"] + [str(f) for f in cfg["frame_ids"]]
            tr = Trainer(MonodepthOptions().parse(argv))
            for k, m in tr.models.items():
                fill_deterministic(m, salt=k + "/")
            tr.epoch = cfg["epoch"]
            inputs, noise, xy = C.step_inputs(cfg)
            with torch.no_grad():
                out, losses = tr.process_batch(dict(inputs), noise=noise, mask_xy=xy)
            fr = []
            for key in ("depth_mvs", "masked_depth", "fused_depth"):
                got = out[key].cpu().numpy().reshape(gold[key].shape)
                fr.append(float((np.abs(got - gold[key]) / np.abs(gold[key]) < 1e-3).mean()))
            d0 = out[("disp", 0)].cpu().numpy()
            drel = float((np.abs(d0 - gold["disp0"]) / np.abs(gold["disp0"])).max())
            lrel = abs(float(losses["loss"]) - float(gold["loss/loss"])) / abs(float(gold["loss/loss"]))
            print("%-10s %-11s %9.4f %9.4f %9.4f %11.2e %11.2e" % (name, pol, fr[0], fr[1], fr[2], drel, lrel), flush=True)


if __name__ == "__main__":
    main()
