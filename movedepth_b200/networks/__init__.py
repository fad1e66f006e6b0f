"""Sub-networks of the hot path, constructor-compatible with the reference's `movedepth.networks`
(movedepth/networks/__init__.py:2-5) and state-dict compatible with its checkpoints."""
from .resnet_encoder import ResnetEncoder, FPN4, reg3d, reg2d
from .depth_decoder import DepthDecoder, UncertNet
from .pose_decoder import PoseDecoder
