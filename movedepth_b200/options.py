"""Command-line flag set, name- and default-compatible with the reference's `MonodepthOptions`
(movedepth/options.py:7-350) so its launch scripts (train_movedepth.sh:16-30) work unchanged.

Only the flags the hot path reads change behaviour here; the rest are
accepted so existing command lines parse.  Extra flags of this build are prefixed `--b200_`.
"""
import argparse
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

# (name, type, default[, choices])
_VALUES = [
    ("data_path", str, os.path.join(_HERE, "kitti_data")), ("log_dir", str, os.path.join(os.path.expanduser("~"), "tmp")),
    ("model_name", str, "mdp"), ("split", str, "eigen_zhou"), ("num_layers", int, 18, [18, 34, 50, 101, 152]),
    ("depth_binning", str, "linear", ["linear", "inverse"]), ("num_depth_bins", int, 16),
    ("ztrans_start_epc", int, 8), ("depth_bin_fac", float, 0.3), ("ssim_lw", float, 0.85),
    ("split1", float, 0.333), ("split2", float, 0.666), ("mask_lw", float, 10), ("photo_conf", float, 0.2),
    ("dataset", str, "kitti"), ("height", int, 192), ("width", int, 640), ("disparity_smoothness", float, 1e-3),
    ("min_depth", float, 0.1), ("max_depth", float, 100.0), ("batch_size", int, 12), ("res_arch", int, 18),
    ("learning_rate", float, 1e-4), ("num_epochs", int, 20), ("scheduler_step_size", int, 15),
    ("pytorch_random_seed", int, None), ("update_range_epoch", int, 0), ("lr_fac", float, 1),
    ("weights_init", str, "pretrained", ["pretrained", "scratch"]), ("num_matching_frames", int, 1),
    ("reg3d_c", int, 16), ("prior_scale", int, 2), ("norm_radius", int, 1), ("schedule_type", str, "inverse"),
    ("iter_stages", int, 4), ("iter_bins", int, 8), ("z_scale", float, 30), ("dist_thres", float, 0),
    ("num_workers", int, 12), ("load_weights_folder", str, None), ("mono_weights_folder", str, None),
    ("log_frequency", int, 250), ("save_frequency", int, 1), ("pred_depth_scale_factor", float, 1),
    ("ext_disp_to_eval", str, None), ("eval_split", str, "eigen"), ("eval_out_dir", str, None),
    ("pixel_thres", float, 1), ("depth_thres", float, 0.1), ("freeze_fuse_epc", int, 0), ("local_rank", int, 0),
]
_LISTS = [
    ("scales", int, [0, 1, 2, 3]), ("frame_ids", int, [0, -1, 1]), ("matching_ids", int, [0, -1]),
    ("casbins", int, [8, 4, 4]), ("casfac", float, [0.5, 0.25, 0.125]), ("casch", int, [8, 4, 4]),
    ("models_to_load", str, ["encoder", "depth", "pose_encoder", "pose, reg3d", "mono_depth", "mono_encoder"]),
]
_SWITCHES = [
    "png", "v1_multiscale", "avg_reprojection", "disable_automasking", "enable_mvs_pose_grad", "no_ssim",
    "use_future_frame", "disable_motion_masking", "disable_edge_masking", "no_matching_augmentation", "group_cor",
    "mvs_norm", "conv3d", "mono_prior", "preconv", "log", "fix_scale", "mvs_cascade", "mvs_raft", "no_cuda",
    "save_intermediate_models", "eval_stereo", "eval_mono", "disable_median_scaling", "save_pred_disps", "no_eval",
    "eval_eigen_to_benchmark", "post_process", "zero_cost_volume", "static_camera", "eval_teacher", "convex_up",
    "load_pose", "mask_mvs_conf", "mask_mvs_dist", "mask_mvs_geo", "mask_mvs_auto", "mvs_smooth_loss", "dcn",
    "train_motion_only", "ddp",
]


class MonodepthOptions:
    def __init__(self):
        p = argparse.ArgumentParser(description="MOVEDepth options (B200-native build)")
        for spec in _VALUES:
            name, typ, default = spec[:3]
            kw = {"choices": spec[3]} if len(spec) > 3 else {}
            p.add_argument("--" + name, type=typ, default=default, **kw)
        for name, typ, default in _LISTS:
            p.add_argument("--" + name, type=typ, nargs="+", default=list(default))
        for name in _SWITCHES:
            p.add_argument("--" + name, action="store_true")
        # torch >= 2 launchers pass --local-rank; accept both spellings
        p.add_argument("--local-rank", dest="local_rank", type=int, default=0)
        # ---- this build only
        p.add_argument("--b200_conv_precision", choices=["fp32", "3xtf32", "tf32"], default="3xtf32",
                       help="conv arithmetic (movedepth_b200/precision.py): fp32 = cuDNN SIMT kernels; 3xtf32 = tensor-core "
                            "convs with a 3-way TF32 operand split in the forward (near-fp32 outputs), TF32 gradients; "
                            "tf32 = PyTorch's default conv policy")
        p.add_argument("--b200_split_backward", action="store_true", help="3xtf32: split dgrad/wgrad as well")
        p.add_argument("--b200_cuda_graph", action="store_true", help="capture the training step in a CUDA graph")
        p.add_argument("--b200_one_stream", action="store_true",
                       help="issue the mono/pose graph and the cost-volume graph on ONE stream (default: two concurrent streams)")
        p.add_argument("--b200_cudnn_benchmark", action="store_true",
                       help="let cuDNN time its algorithms per layer shape during the warm-up steps (torch.backends.cudnn.benchmark; "
                            "the reference's train.py:19-20 pins deterministic=True, benchmark=False)")
        p.add_argument("--b200_synthetic", action="store_true", help="train on synthetic KITTI-shape tensors")
        self.parser = p

    def parse(self, args=None):
        self.options = self.parser.parse_args(args)
        return self.options


# the reference's train.py imports this (non-existent there) name -- provide it
MovedepthOptions = MonodepthOptions
