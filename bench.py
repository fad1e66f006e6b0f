#!/usr/bin/env python
"""Benchmark of the MOVEDepth dense hot path (BASELINE.json metric: training frames/sec at
192x640, D=96; cost-volume kernel HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {2,3,5}] [--impl reference]

Own arm: one process per GPU (launched by torchrun for N>1).  A step = Trainer.process_batch +
backward + fused Adam on one synthetic batch of the chosen BASELINE config (default configs[1]:
ResNet18, 2-frame, 192x640, D=96, batch 6 per GPU), with the velocity-guided hypothesis range the
north_star names (epoch > ztrans_start_epc).  Prints ONE JSON line on rank 0.
Reference arm (`--impl reference`): the UNMODIFIED reference (`baseline/_ref`, installed by
baseline/install_reference.sh) driven through its own Trainer on the host cores, same config and
batch; when that install is absent, the CPU oracle port (oracle/step.py).
"""
import argparse
import contextlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, G = 32, 16            # matching-feature channels, correlation groups
CONFIGS = {
    2: dict(arch=18, frame_ids=[0, -1], H=192, W=640, D=96, B=6, tag="BASELINE configs[1]", desc="ResNet18 2-frame 192x640 D=96 batch 6/GPU"),
    3: dict(arch=50, frame_ids=[0, -1, 1], H=192, W=640, D=96, B=6, tag="BASELINE configs[2]", desc="ResNet50 3-frame 192x640 D=96 batch 6/GPU"),
    5: dict(arch=50, frame_ids=[0, -1, 1], H=320, W=1024, D=128, B=4, tag="BASELINE configs[4]", desc="ResNet50 3-frame 320x1024 D=128 batch 4/GPU"),
}


def metric_name(cfg):
    return "training frames/sec (%dx%d, D=%d)" % (cfg["H"], cfg["W"], cfg["D"])


def workload(cfg):
    return "%s, fwd+bwd+Adam (%s)" % (cfg["desc"], cfg["tag"])


def shared_config(cfg, args):
    """The `config` object both arms print (same workload, batch, schedule and data on both sides)."""
    return {"workload": workload(cfg), "batch_per_step_per_gpu": cfg["B"],
            "schedule": "fixed range (epoch 0)" if args.fixed_range else "velocity-guided range (epoch 9 > ztrans_start_epc)",
            "images": "U[0,1) white noise (SURVEY 8d)" if args.noise else "band-limited noise, source frames shifted 2 px per frame index",
            "l2": "GPU arm: a 256 MiB buffer is written between timed steps (L2 flush); per-step activations also exceed the 126 MB L2"}


def costvol_bytes(cfg):
    """Algorithmic bytes of one fused cost-volume launch: ref + src + prior in, grouped volume out
    = 4*B*h*w*(2C + 1 + D*G)  (SURVEY section 8d)."""
    return 4 * cfg["B"] * (cfg["H"] // 4) * (cfg["W"] // 4) * (2 * C + 1 + cfg["D"] * G)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.05)
        except Exception as e:                       # NVML missing: report that rather than inventing clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0}, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram bytes per launch of the cost-volume forward from the committed `ncu --set full` capture, or None."""
    p = os.path.join(ROOT, "profiles", "costvol_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("dram_bytes_per_launch"), d.get("source")
    return None, None


# ---------------------------------------------------------------------------------------------- reference arm
def oracle_options(cfg, batch):
    from oracle.step import default_options
    return default_options(height=cfg["H"], width=cfg["W"], num_depth_bins=cfg["D"], batch_size=batch, frame_ids=list(cfg["frame_ids"]),
                           matching_ids=[0, -1], res_arch=cfg["arch"], learning_rate=2e-4)


class _PortStep:
    def __init__(self, cfg, batch, make_inputs, epoch):
        import torch
        from oracle.step import OracleStep
        torch.set_num_threads(os.cpu_count() or 1)
        opt = oracle_options(cfg, batch)
        torch.manual_seed(0)
        self.st, self.inputs, self.epoch, self.batch = OracleStep(opt), make_inputs(opt, batch), epoch, batch
        self.threads = torch.get_num_threads()

    def step(self):
        self.st.train_step(dict(self.inputs), epoch=self.epoch)


def time_cpu(cfg, args, steps, warmup, batch, budget_s):
    """frames/s of the reference's training step on the host cores: the unmodified reference from baseline/_ref when it is
    installed (kind "reference"), else the oracle port (kind "port").  The step runs the workload's full batch; only when
    the first (warm-up) step shows that `steps + warmup` of them cannot finish within `budget_s` is the per-step sample
    bounded to fewer frames (stated in the returned `sample`).  Returns (fps, sec/step, threads, kind, frames per step)."""
    from oracle.step import synthetic_inputs
    from baseline import reference_runner as RR
    epoch = 0 if args.fixed_range else 9
    make = lambda opt, b: synthetic_inputs(opt, b, seed=1, smooth=not args.noise)
    use_ref = RR.available() and not args.port
    rcfg = dict(height=cfg["H"], width=cfg["W"], num_depth_bins=cfg["D"], res_arch=cfg["arch"], frame_ids=cfg["frame_ids"], epoch=epoch)
    build = (lambda b: RR.ReferenceStep(rcfg, b, make)) if use_ref else (lambda b: _PortStep(cfg, b, make, epoch))
    with contextlib.redirect_stdout(sys.stderr):              # the reference prints its banner on stdout
        runner = build(batch)
        t0 = time.perf_counter()
        runner.step()                                         # first warm-up step, timed to size the sample
        t1 = time.perf_counter() - t0
        if t1 * (steps + warmup) > budget_s and batch > 1:
            batch = max(1, int(batch * budget_s / (t1 * (steps + warmup))))
            runner = build(batch)
            runner.step()
        for _ in range(warmup - 1):
            runner.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            runner.step()
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, runner.threads, "reference" if use_ref else "port", batch


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    warm = max(1, args.warmup)
    fps, sec, cores, kind, frames = time_cpu(cfg, args, args.steps, warm, cfg["B"], budget_s=args.cpu_budget)
    sample = "%d frames per step (%s), %d timed steps after %d warm-up, torch CPU with %d threads" % (
        frames, "the workload's batch" if frames == cfg["B"] else "bounded: %d steps of the full batch of %d exceed %d s on this host"
        % (args.steps + warm, cfg["B"], args.cpu_budget), args.steps, warm, cores)
    print(json.dumps({
        "impl": "reference", "metric": metric_name(cfg), "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": shared_config(cfg, args),
        "execution": {"host": "CPU, torch %d threads" % cores,
                      "what": "unmodified reference Trainer (baseline/_ref) on the host cores" if kind == "reference"
                      else "oracle port of the reference step (oracle/step.py); baseline/_ref is not installed"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------- own arm
def shutdown(tr, world):
    """Orderly multi-rank exit: drop the captured graphs (they hold NCCL work), meet at a barrier, then leave without
    running interpreter teardown (a captured communicator can block in its destructor)."""
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        tr._graphs.clear()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


class StepModel:
    """Byte / FLOP model of one training step for the step-level roofline: forward hooks on every
    convolution record its logical FLOPs and its input + output bytes during one eager step.
      FLOPs:  forward + data gradient + weight gradient = 3 x forward (logical fp32 FLOPs); under the 3xTF32 policy the
              forward executes 3 tensor-core products per logical one -> executed = 5/3 x logical;
      bytes:  the minimal conv traffic with every BatchNorm / activation / pad / split fused away: forward reads the input
              and writes the output once, the two gradient passes read (gy, x) and write gx -> 3 x (in + out) x 4 B, plus the
              cost-volume kernels' algorithmic bytes (4 passes per step: 2 forward, 2 backward)."""

    def __init__(self, models):
        import torch.nn as nn
        from movedepth_b200 import precision as PR
        self.flops_fwd, self.bytes_fwd, self.handles = 0, 0, []
        for m in models.values():
            for mod in m.modules():
                if isinstance(mod, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose3d)):
                    self.handles.append(mod.register_forward_hook(self._hook))
        PR.conv_recorder = []            # convolutions issued without going through their module's __call__ (fused paths)

    def _hook(self, mod, inp, out):
        import torch.nn as nn
        x = inp[0]
        k = 1
        for v in mod.kernel_size:
            k *= v
        if isinstance(mod, nn.ConvTranspose3d):
            self.flops_fwd += 2 * x.numel() * (mod.out_channels // mod.groups) * k
        else:
            self.flops_fwd += 2 * out.numel() * (mod.in_channels // mod.groups) * k
        self.bytes_fwd += 4 * (x.numel() + out.numel())

    def remove(self):
        from movedepth_b200 import precision as PR
        for h in self.handles:
            h.remove()
        for wshape, xshape, yshape, transposed in (PR.conv_recorder or []):
            k, numel = 1, lambda s: int(__import__("math").prod(s))
            for v in wshape[2:]:
                k *= v
            if transposed:
                self.flops_fwd += 2 * numel(xshape) * wshape[1] * k
            else:
                self.flops_fwd += 2 * numel(yshape) * wshape[1] * k
            self.bytes_fwd += 4 * (numel(xshape) + numel(yshape))
        PR.conv_recorder = None

    def summary(self, cfg, precision, ms_per_step, pk):
        flops = 3 * self.flops_fwd
        executed = flops * 5 / 3 if precision == "3xtf32" else flops
        nbytes = 3 * self.bytes_fwd + 4 * costvol_bytes(cfg)
        tf32_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / 2          # dense TF32 = half the bf16 rate
        t_hbm, t_tensor = nbytes / (pk["hbm_gbs"] * 1e9) * 1e3, executed / (tf32_peak * 1e12) * 1e3
        return {"flops_logical": flops, "flops_executed_tensor": executed, "bytes_min": nbytes, "t_hbm_ms": t_hbm, "t_tensor_ms": t_tensor,
                "bound": "hbm" if t_hbm >= t_tensor else "tensor", "frac": max(t_hbm, t_tensor) / ms_per_step,
                "tf32_peak_tflops": tf32_peak, "hbm_peak_gbs": pk["hbm_gbs"],
                "model": "convs: 3 x forward FLOPs (x5/3 executed under 3xTF32), 3 x (in+out) x 4 B with all elementwise work fused away; "
                         "+ 4 cost-volume passes; frac = max(t_hbm, t_tensor) / measured step time"}


def run_own(args, cfg):
    import torch
    import torch.distributed as dist
    from movedepth_b200 import ops
    from movedepth_b200.options import MonodepthOptions
    from movedepth_b200.trainer import Trainer, SyntheticKITTI

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    BATCH = cfg["B"]
    argv = ["--height", str(cfg["H"]), "--width", str(cfg["W"]), "--num_depth_bins", str(cfg["D"]), "--batch_size", str(BATCH),
            "--frame_ids"] + [str(f) for f in cfg["frame_ids"]] + [
            "--matching_ids", "0", "-1", "--res_arch", str(cfg["arch"]), "--weights_init", "scratch", "--prior_scale", "2", "--convex_up",
            "--learning_rate", "2e-4", "--b200_conv_precision", args.precision, "--log_dir", "/tmp/mvd_bench"]
    if world > 1:
        argv.append("--ddp")
    if not args.no_graph:
        argv.append("--b200_cuda_graph")
    if not args.no_cudnn_benchmark:
        argv.append("--b200_cudnn_benchmark")
    opt = MonodepthOptions().parse(argv)
    torch.manual_seed(0)
    tr = Trainer(opt)
    tr.epoch = 0 if args.fixed_range else 9      # epoch > ztrans_start_epc: the velocity-guided hypothesis range
    # Synthetic KITTI-shape batches.  Default: band-limited noise images whose source frames are the reference frame shifted by
    # 2 px per frame index (image-like statistics and a consistent inter-frame motion, so the self-supervised losses have a
    # signal and the predicted pose / depth stay in the regime the cost-volume kernel is designed for).  `--noise` selects the
    # U[0,1) white-noise images of SURVEY section 8(d): the dense kernels are data independent (same step time), but the
    # networks then drift to arbitrary poses / saturated depths within ~10 steps and the cost volume's gather locality with them.
    host_batches = list(SyntheticKITTI(opt, BATCH, 4, seed=1 + rank, smooth=not args.noise))
    dev_batches = [{k: v.cuda(non_blocking=True) for k, v in b.items()} for b in host_batches]
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, read_loss, graph_pairs=None, cv=None):
        """K steps bracketed by barrier + synchronize, device-timed.  With `graph_pairs` (the cost-volume timing-event nodes
        inside the replayed graph) every step ends with a stream synchronize so the pair can be read before the next replay
        overwrites it -- the kernel is then timed on exactly the steps `value` is measured on (costs ~0.1 % of a step)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            flush.fill_(i & 1)                                            # L2 flush between iterations
            _, losses = tr.train_step(batches[i % len(batches)])
            if read_loss:
                float(losses["loss"].detach())                            # device -> host read of the step result
            if graph_pairs is not None:
                torch.cuda.current_stream().synchronize()
                cv.extend(a.elapsed_time(b) for a, b in graph_pairs)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    ops.costvol_events = []                      # the cost-volume forward records (start, stop) events around its launch
    # graph mode: train_step itself runs GRAPH_WARMUP eager steps and captures on the next call
    warm = max(3, args.warmup) if args.no_graph else max(args.warmup, tr.GRAPH_WARMUP + 1)
    model = StepModel(tr.models)
    for i in range(warm):
        tr.train_step(dev_batches[i % 4])
        if i == 0:
            model.remove()                       # one eager step is enough to see every convolution's shape
    torch.cuda.synchronize()
    graph_pairs = None if args.no_graph else ops.costvol_events[-2:]             # the event nodes inside the captured graph
    sampler = ClockSampler(local)
    sampler.start()
    ops.costvol_events = []
    n0 = ops.launch_counter["n"]
    if args.ncu_range:                       # `ncu --profile-from-start off`: capture exactly the timed steps
        torch.cuda.profiler.start()
    cv = []
    ms = timed(dev_batches, args.steps, read_loss=False, graph_pairs=graph_pairs, cv=cv)
    if args.ncu_range:
        torch.cuda.profiler.stop()
    launches = ops.launch_counter["n"] - n0
    if graph_pairs is None:
        cv = [a.elapsed_time(b) for a, b in ops.costvol_events]
    ops.costvol_events = None
    if args.verbose and rank == 0:               # what did K1 see?  (pose / prior statistics of the last step)
        out, _ = tr.train_step(dev_batches[0])
        T = out[("cam_T_cam", 0, -1)].detach()
        pr = out["depth_prior"].detach()
        print("K1 inputs: |t| per item %s, tz %s, prior min/median/max %.3f/%.3f/%.3f; K1 us per launch %s" % (
            [round(float(v), 4) for v in T[:, :3, 3].norm(dim=1)], [round(float(v), 4) for v in T[:, 2, 3]],
            float(pr.min()), float(pr.median()), float(pr.max()), [round(v * 1e3) for v in cv]), file=sys.stderr)
    ms_e2e = timed(host_batches, args.steps, read_loss=True)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        shutdown(tr, world)
        return
    frames = BATCH * world * args.steps
    pk, pk_src = peaks()
    cv_ms = sum(cv) / len(cv)
    cvb = costvol_bytes(cfg)
    achieved = cvb / (cv_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic() if args.config == 2 else (None, None)
    grads = {"3xtf32": "TF32 gradients (single pass, operands truncated to 10 mantissa bits by the tensor core)",
             "fp32": "fp32 gradients", "tf32": "TF32 gradients"}[args.precision]
    line = {
        "metric": metric_name(cfg), "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"3xtf32": "f32 storage; 3xTF32 forward (fp32-class), TF32 gradients", "fp32": "f32", "tf32": "tf32"}[args.precision],
        "data": "synthetic (U[0,1) white noise, SURVEY 8d)" if args.noise else "synthetic (band-limited noise images, source frames shifted 2 px)",
        "config": shared_config(cfg, args),
        "execution": {"precision": {"3xtf32": "convs on tensor cores with a 3-way TF32 operand split in the forward (near-fp32 outputs, "
                                    "movedepth_b200/precision.py); " + grads,
                                    "fp32": "fp32 everywhere (cuDNN SIMT convs)", "tf32": "cuDNN TF32 everywhere"}[args.precision],
                      "cudnn": "library convs: algorithm per layer shape timed by cuDNN during the warm-up steps (cudnn.benchmark)"
                               if not args.no_cudnn_benchmark else "library convs: cuDNN heuristics",
                      "launch": "eager launches" if args.no_graph else "forward+backward(+all-reduce) replayed as one CUDA graph; Adam kernels after it",
                      "parallelism": "dp%d, flat-arena gradient all-reduce over NCCL" % world},
        "roofline": {"kernel": "costvol_grouped_fwd_kernel", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk_src,
                     "frac_of_nominal_8000": achieved / 8000.0,
                     "algorithmic_bytes": cvb, "avg_launch_us": cv_ms * 1e3, "launches_timed": len(cv),
                     "launch_us_min_median_max": [min(cv) * 1e3, statistics.median(cv) * 1e3, max(cv) * 1e3],
                     "note": "timed inside the training step (event nodes in the replayed graph) on the features / prior / pose the "
                             "networks produce at that step; velocity-guided range: the per-item scale is depth_bin_fac*z_scale*T[b,2,3]"},
        "step_roofline": model.summary(cfg, args.precision, ms / args.steps, pk),
        "clocks": sampler.summary(),
        "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
    }
    if world == 1 and not args.no_cpu_baseline:
        fps, sec, cores, kind, frames = time_cpu(cfg, args, 2, 1, BATCH, budget_s=45)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                                "sample": "%d frames per step (workload batch %d), 2 timed steps after 1 warm-up" % (frames, BATCH)}
    print(json.dumps(line), flush=True)
    shutdown(tr, world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config number (1-based): 2, 3 or 5")
    ap.add_argument("--precision", default="3xtf32", choices=["3xtf32", "fp32", "tf32"])
    ap.add_argument("--fixed_range", action="store_true", help="epoch-0 fixed hypothesis range instead of the velocity-guided one")
    ap.add_argument("--velocity", action="store_true", help="(default now) velocity-guided hypothesis range; kept for old command lines")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--cpu_budget", type=int, default=600, help="reference arm: seconds the K+W host steps may take before the per-step sample is bounded")
    ap.add_argument("--port", action="store_true", help="reference arm / cpu_baseline: time the oracle port even when baseline/_ref exists")
    ap.add_argument("--noise", action="store_true", help="U[0,1) white-noise images (SURVEY 8d) instead of band-limited ones")
    ap.add_argument("--verbose", action="store_true", help="print the pose / prior statistics the cost-volume kernel saw")
    ap.add_argument("--no_graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no_cudnn_benchmark", action="store_true", help="cuDNN heuristics instead of timed algorithm selection for the library convs")
    ap.add_argument("--ncu_range", action="store_true", help="cudaProfilerStart/Stop around the device-timed steps")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_own(args, cfg)


if __name__ == "__main__":
    main()
