"""Runs the UNMODIFIED reference (`baseline/_ref/movedepth`, installed by baseline/install_reference.sh) through its own
public API on the host CPU: `Trainer(opts)` -> `process_batch` -> `backward` -> `model_optimizer.step()`
(movedepth/trainer.py:269-272), on synthetic KITTI-shape item dicts.  Used by `bench.py --impl reference` and the
`cpu_baseline` leg only -- never by the product.

Environment shims (none touches the arithmetic):
  * stub modules for imports that are absent here and unused on this path: tensorboardX, matplotlib, skimage, pykitti;
  * `PIL.Image.ANTIALIAS` alias (removed in Pillow 10; referenced at import time by datasets/mono_dataset.py:56);
  * `torch.cuda.set_device` -> no-op (trainer.py:47 calls it even with --no_cuda);
  * `UncertNet.forward`: the in-place residual `out += x` (networks/depth_decoder.py:390) written out of place -- same
    values; torch >= 2 refuses to back-propagate through the in-place form;
  * options are parsed with `MonodepthOptions` directly (train.py:5 imports a name that does not exist).
"""
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "movedepth", "trainer.py"))


def _import_reference():
    import torch

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    class _Writer:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, _):
            return lambda *a, **k: None

    stub("tensorboardX", SummaryWriter=_Writer)
    plt = stub("matplotlib.pyplot", get_cmap=lambda *a, **k: None)
    stub("matplotlib", pyplot=plt)
    tr = stub("skimage.transform")
    stub("skimage", transform=tr)
    stub("pykitti")
    from PIL import Image
    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    torch.cuda.set_device = lambda *_: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import movedepth.networks as rn
    import movedepth.trainer as rt
    from movedepth.options import MonodepthOptions

    def uncert_forward(self, x):
        out = self.conv2(self.conv1(x))
        out = out + x
        return torch.sigmoid(self.head_convs(out))
    rn.UncertNet.forward = uncert_forward
    return rt, MonodepthOptions


class ReferenceStep:
    """The reference's own training step on the host cores: Trainer(opts) -> process_batch -> backward -> Adam.step.
    cfg: dict(height, width, num_depth_bins, res_arch, frame_ids, epoch); make_inputs(opt, batch) -> item dict."""

    def __init__(self, cfg, batch, make_inputs, threads=None):
        import torch
        torch.set_num_threads(threads or os.cpu_count() or 1)
        rt, Options = _import_reference()
        argv = ["--no_cuda", "--weights_init", "scratch", "--num_workers", "0", "--data_path", "/nonexistent", "--png",
                "--log_dir", "/tmp/mvd_reference", "--prior_scale", "2", "--convex_up", "--learning_rate", "2e-4",
                "--height", str(cfg["height"]), "--width", str(cfg["width"]), "--num_depth_bins", str(cfg["num_depth_bins"]),
                "--batch_size", str(batch), "--res_arch", str(cfg["res_arch"]),
                "--frame_ids"] + [str(f) for f in cfg["frame_ids"]]
        opt = Options().parser.parse_args(argv)
        torch.manual_seed(0)
        self.trainer = rt.Trainer(opt)
        self.trainer.set_train()
        self.trainer.epoch, self.trainer.step = cfg["epoch"], 0
        self.inputs = make_inputs(opt, batch)
        self.batch = batch
        self.threads = torch.get_num_threads()

    def step(self):
        tr = self.trainer
        outputs, losses = tr.process_batch(dict(self.inputs), is_train=True)
        tr.model_optimizer.zero_grad()
        losses["loss"].backward()
        tr.model_optimizer.step()
        return float(losses["loss"].detach())
