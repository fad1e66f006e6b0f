"""Where does a whole-step parity case drift?  Runs the CPU oracle step and the GPU Trainer on the same
deterministic weights / inputs and prints max abs / rel error of the intermediates, per cuDNN setting.
    python tools/diag_parity.py [case] """
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as C  # noqa: E402
from _weights import fill_deterministic  # noqa: E402
from oracle.step import OracleStep  # noqa: E402
from movedepth_b200.options import MonodepthOptions  # noqa: E402
from movedepth_b200.trainer import Trainer  # noqa: E402


def err(a, b):
    a = a.detach().float().cpu().reshape(b.shape)
    b = b.detach().float().cpu()
    d = (a - b).abs()
    return "max|d| %.3e  max|b| %.3e  rel(max) %.3e" % (float(d.max()), float(b.abs().max()), float(d.max() / b.abs().max()))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "r50_3f"
    cfg = C.STEP_CASES[name]
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "step_%s.npz" % name)))
    opt = C.step_options(cfg)
    st = OracleStep(opt)
    for k, m in st.models.items():
        fill_deterministic(m, salt=k + "/")
    inputs, noise, xy = C.step_inputs(cfg)
    feats_o = []
    st.models["mvs_encoder"].register_forward_hook(lambda m, a, o: feats_o.append(o[0].detach()))
    out_o, _ = st.process_batch(dict(inputs), epoch=cfg["epoch"], noise=[n.clone() for n in noise], mask_xy=xy)
    print("oracle vs golden cost_volume:", err(out_o["cost_volume"], torch.from_numpy(gold["cost_volume"])))

    for bench in (True, False):
        argv = ["--height", str(cfg["H"]), "--width", str(cfg["W"]), "--num_depth_bins", str(cfg["D"]), "--batch_size",
                str(cfg["B"]), "--res_arch", str(cfg.get("arch", 18)), "--weights_init", "scratch", "--convex_up",
                "--b200_conv_precision", "fp32", "--log_dir", "/tmp/mvd_diag", "--frame_ids"] + [str(f) for f in cfg["frame_ids"]]
        tr = Trainer(MonodepthOptions().parse(argv))
        torch.backends.cudnn.benchmark = bench
        for k, m in tr.models.items():
            fill_deterministic(m, salt=k + "/")
        tr.epoch = cfg["epoch"]
        feats_g = []
        tr.models["mvs_encoder"].register_forward_hook(lambda m, a, o: feats_g.append(o[0].detach()))
        with torch.no_grad():
            out, _ = tr.process_batch(dict(inputs), noise=[n.clone() for n in noise], mask_xy=xy)
        print("---- cudnn.benchmark =", bench)
        for i, (a, b) in enumerate(zip(feats_g, feats_o)):
            print("  FPN4 call %d          :" % i, err(a, b))
        for f in cfg["frame_ids"][1:]:
            print("  cam_T_cam %2d         :" % f, err(out[("cam_T_cam", 0, f)], out_o["cam_T_cam", 0, f]))
        print("  disp2                :", err(out[("disp", 2)], out_o["disp", 2]))
        hyps_g = out["depth_prior"] * out["hypothesis_ratio"][:, :, None, None]
        print("  hypotheses           :", err(hyps_g, out_o["depth_hypotheses"]))
        print("  cost_volume          :", err(out["cost_volume"].permute(0, 2, 1, 3, 4), out_o["cost_volume"]))
        print("  cost_volume vs gold  :", err(out["cost_volume"].permute(0, 2, 1, 3, 4), torch.from_numpy(gold["cost_volume"])))
        # same kernel on the ORACLE's features / hypotheses: isolates K1 from upstream drift
        from movedepth_b200 import ops
        dev = tr.device
        vol = ops.costvol_grouped(feats_o[0].to(dev), feats_o[1].to(dev), inputs["K", 2].to(dev), inputs["inv_K", 2].to(dev),
                                  inputs_pose(out_o, cfg).to(dev), hyps=out_o["depth_hypotheses"].to(dev))
        print("  K1 on oracle inputs  :", err(vol.permute(0, 2, 1, 3, 4), out_o["cost_volume"]))


def inputs_pose(out_o, cfg):
    return out_o["cam_T_cam", 0, -1].detach()


if __name__ == "__main__":
    main()
