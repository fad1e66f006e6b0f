"""Mono disparity decoder and the entropy -> trust-mono-mask head.

State-dict keys follow the reference (`decoder.{0..13}.conv.conv.*` / `decoder.{10..13}.conv.*`;
`conv{1,2}.{0,1}.*`, `head_convs.weight`), SURVEY.md section 4.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import precision as PR
from ..layers import ConvBlock, Conv3x3, upsample
from .resnet_encoder import ConvBnReLUSeq


fused = True             # set False to run the decoder as separate tensor ops (A/B measurements, parity isolation)


class DepthDecoder(nn.Module):
    """U-Net decoder over the 5 encoder maps -> sigmoid disparity at `scales`.
    Reference: movedepth/networks/depth_decoder.py:10-101 in the configuration the trainer uses
    (trainer.py:74-75: skips on, one output channel; the ddv / mono_conf / match_conv variants
    are never enabled there and are not built)."""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True, **unused):
        super().__init__()
        self.num_output_channels, self.use_skips, self.scales = num_output_channels, use_skips, list(scales)
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array([16, 32, 64, 128, 256])
        enc = [int(c) for c in num_ch_enc]
        dec = [int(c) for c in self.num_ch_dec]
        blocks, self._slot = [], {}

        def add(key, module):
            self._slot[key] = len(blocks)
            blocks.append(module)

        for i in range(4, -1, -1):                       # same insertion order as the reference's OrderedDict
            add(("upconv", i, 0), ConvBlock(enc[-1] if i == 4 else dec[i + 1], dec[i]))
            add(("upconv", i, 1), ConvBlock(dec[i] + (enc[i - 1] if (use_skips and i > 0) else 0), dec[i]))
        for s in self.scales:
            add(("dispconv", s), Conv3x3(dec[s], num_output_channels))
        self.decoder = nn.ModuleList(blocks)

    def _m(self, *key):
        return self.decoder[self._slot[key]]

    def _forward_fused(self, input_features):
        """CUDA path: between two convolutions everything (bias, ELU, nearest x2, skip concatenation, reflection padding and
        the 3xTF32 operand split) is ONE `decoder_prep` kernel; the convolutions themselves see pre-padded inputs."""
        from .. import ops
        want = PR.get_policy() == "3xtf32"
        self.outputs = {}
        xp, x3 = ops.decoder_prep(input_features[-1], None, None, act=False, up=1, want_split=want)
        for i in range(4, -1, -1):
            c0 = self._m("upconv", i, 0).conv.conv
            z = PR.prepared_conv(xp, x3, c0.weight)
            skip = input_features[i - 1] if (self.use_skips and i > 0) else None
            c1 = self._m("upconv", i, 1).conv.conv
            xp, x3 = ops.decoder_prep(z, c0.bias, skip, act=True, up=2, want_split=want and not PR.prepared_is_small(c1.weight))
            z = PR.prepared_conv(xp, x3, c1.weight)
            if i in self.scales or i > 0:                # this padded activation feeds dispconv(i) and upconv(i-1, 0)
                direct = i == 0 and i in self.scales and PR.prepared_is_small(self._m("dispconv", i).conv.weight)
                xp, x3 = ops.decoder_prep(z, c1.bias, None, act=True, up=1, want_split=want and not direct)
            if i in self.scales:
                dc = self._m("dispconv", i).conv
                self.outputs[("disp", i)] = torch.sigmoid(PR.prepared_conv(xp, x3, dc.weight) + dc.bias.view(1, -1, 1, 1))
        return self.outputs

    def forward(self, input_features, **unused):
        if input_features[-1].is_cuda and input_features[-1].dtype == torch.float32 and fused:
            return self._forward_fused(input_features)
        self.outputs = {}
        x = input_features[-1]
        for i in range(4, -1, -1):
            x = upsample(self._m("upconv", i, 0)(x))
            if self.use_skips and i > 0:
                x = torch.cat([x, input_features[i - 1]], 1)
            x = self._m("upconv", i, 1)(x)
            if i in self.scales:
                self.outputs[("disp", i)] = torch.sigmoid(self._m("dispconv", i)(x))
        return self.outputs


class UncertNet(nn.Module):
    """Entropy map [B,1,h,w] -> trust-mono mask in (0,1).  Reference: depth_decoder.py:371-393.
    The 1-channel input is added (broadcast) to the 8-channel feature before the head; written
    out of place (the reference's in-place add breaks autograd on current torch)."""

    def __init__(self):
        super().__init__()
        self.conv1 = ConvBnReLUSeq(PR.Conv2d(1, 8, 3, 1, 1, bias=False), nn.BatchNorm2d(8), nn.ReLU(inplace=True))
        self.conv2 = ConvBnReLUSeq(PR.Conv2d(8, 8, 3, 1, 1, bias=False), nn.BatchNorm2d(8), nn.ReLU(inplace=True))
        self.head_convs = PR.Conv2d(8, 1, 3, 1, 1, bias=False)

    def forward(self, x):
        y = self.conv2(self.conv1(x)) + x
        return torch.sigmoid(self.head_convs(y))
