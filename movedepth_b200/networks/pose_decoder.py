"""Pose head.  State-dict keys `net.{0..3}.{weight,bias}` as in the reference."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class PoseDecoder(nn.Module):
    """squeeze 1x1 -> 2 x conv3x3 -> 1x1 -> spatial mean -> x0.01 -> (axisangle, translation),
    each [B, num_frames, 1, 3].  Reference: movedepth/networks/pose_decoder.py:8-48."""

    def __init__(self, num_ch_enc, num_input_features, num_frames_to_predict_for=None, stride=1):
        super().__init__()
        if num_frames_to_predict_for is None:
            num_frames_to_predict_for = num_input_features - 1
        self.num_ch_enc, self.num_input_features = num_ch_enc, num_input_features
        self.num_frames_to_predict_for = num_frames_to_predict_for
        self.net = nn.ModuleList([
            nn.Conv2d(int(num_ch_enc[-1]), 256, 1),
            nn.Conv2d(num_input_features * 256, 256, 3, stride, 1),
            nn.Conv2d(256, 256, 3, stride, 1),
            nn.Conv2d(256, 6 * num_frames_to_predict_for, 1)])

    def forward(self, input_features):
        x = torch.cat([F.relu(self.net[0](f[-1])) for f in input_features], 1)
        x = F.relu(self.net[1](x))
        x = F.relu(self.net[2](x))
        x = self.net[3](x).mean(3).mean(2)
        x = 0.01 * x.view(-1, self.num_frames_to_predict_for, 1, 6)
        return x[..., :3], x[..., 3:]
