"""Oracle restatement of the sub-networks on the hot path (plain torch, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Module/attribute names are chosen so
that `state_dict()` keys equal the reference's (checkpoint key layout is the
compatibility contract); that is also what lets tests load one set of weights into the
reference, the oracle and the CUDA product.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.models as tvm


# ------------------------------------------------------------------ encoders
class ResnetEncoder(nn.Module):
    """movedepth/networks/resnet_encoder.py:74-121 (+21-71 for the multi-image stem).

    torchvision ResNet-18/50 trunk without avgpool/fc; input normalised by (x-0.45)/0.225;
    returns the 5 feature maps.  `num_input_images`>1 widens conv1 to 3*n channels and
    re-initialises convs with kaiming_normal(fan_out) as the reference does."""

    def __init__(self, num_layers, pretrained=False, num_input_images=1):
        super().__init__()
        assert not pretrained, "oracle runs --weights_init scratch only"
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        ctor = {18: tvm.resnet18, 50: tvm.resnet50}[num_layers]
        net = ctor(weights=None)
        if num_input_images > 1:
            net.conv1 = nn.Conv2d(3 * num_input_images, 64, 7, 2, 3, bias=False)
            for m in net.modules():
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                elif isinstance(m, nn.BatchNorm2d):
                    nn.init.constant_(m.weight, 1)
                    nn.init.constant_(m.bias, 0)
        del net.fc, net.avgpool
        self.encoder = net
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4

    def forward(self, image):
        e = self.encoder
        x = (image - 0.45) / 0.225
        f0 = e.relu(e.bn1(e.conv1(x)))
        f1 = e.layer1(e.maxpool(f0))
        f2 = e.layer2(f1)
        f3 = e.layer3(f2)
        f4 = e.layer4(f3)
        return [f0, f1, f2, f3, f4]


class _ConvBnRelu2d(nn.Module):
    """movedepth/networks/resnet_encoder.py:453-475 (`Conv2d`: conv(no bias)+BN+ReLU)."""

    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(cout, momentum=0.1)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class FPN4(nn.Module):
    """movedepth/networks/resnet_encoder.py:311-391 -- matching-feature pyramid.

    4 conv stages (8,16,32,64 ch for base 8), top-down bilinear(align_corners=True) x2 +
    1x1 lateral down to `scale` (2 -> 1/4 res), 3x3 `out` conv -> base*4 channels.
    Returns (matching feature, the stage feature at that resolution)."""

    def __init__(self, base_channels, scale=0, dcn=False):
        super().__init__()
        assert not dcn, "deformable conv is out of scope (external extension, SURVEY §2.2)"
        b = base_channels
        self.scale = scale
        C = _ConvBnRelu2d
        self.conv0 = nn.Sequential(C(3, b, 3, 1, 1), C(b, b, 3, 1, 1))
        self.conv1 = nn.Sequential(C(b, 2 * b, 5, 2, 2), C(2 * b, 2 * b, 3, 1, 1), C(2 * b, 2 * b, 3, 1, 1))
        self.conv2 = nn.Sequential(C(2 * b, 4 * b, 5, 2, 2), C(4 * b, 4 * b, 3, 1, 1), C(4 * b, 4 * b, 3, 1, 1))
        self.conv3 = nn.Sequential(C(4 * b, 8 * b, 5, 2, 2), C(8 * b, 8 * b, 3, 1, 1), C(8 * b, 8 * b, 3, 1, 1))
        fc = 8 * b
        if scale < 3:
            self.inner1 = nn.Conv2d(4 * b, fc, 1, bias=True)
        if scale < 2:
            self.inner2 = nn.Conv2d(2 * b, fc, 1, bias=True)
        if scale < 1:
            self.inner3 = nn.Conv2d(b, fc, 1, bias=True)
        if scale == 3:
            self.out = nn.Conv2d(fc, 8 * b, 1, bias=False)
        else:
            self.out = nn.Conv2d(fc, b * 2 ** scale, 3, padding=1, bias=False)

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        c3 = self.conv3(c2)
        up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
        feat = c3
        if self.scale < 3:
            feat = up(feat) + self.inner1(c2)
        if self.scale < 2:
            feat = up(feat) + self.inner2(c1)
        if self.scale < 1:
            feat = up(feat) + self.inner3(c0)
        return self.out(feat), [c0, c1, c2, c3][self.scale]


# ------------------------------------------------------------------ 3-D regulariser
class _ConvBnRelu3d(nn.Module):
    """movedepth/networks/resnet_encoder.py:175-182."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn = nn.BatchNorm3d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


def _deconv3d(cin, cout):
    return nn.Sequential(
        nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False),
        nn.BatchNorm3d(cout), nn.ReLU())


class Reg3d(nn.Module):
    """movedepth/networks/resnet_encoder.py:227-280 (`reg3d`, down_size=3 as the trainer builds it).

    Input [B,D,G,h,w] (permuted to [B,G,D,h,w]); 3-level 3-D U-Net with additive skips;
    `prob` 3x3x3 conv to one channel -> logits [B,D,h,w]."""

    def __init__(self, in_channels, base_channels, down_size=3):
        super().__init__()
        assert down_size == 3
        b = base_channels
        self.conv0 = _ConvBnRelu3d(in_channels, b)
        self.conv1 = _ConvBnRelu3d(b, 2 * b, 2)
        self.conv2 = _ConvBnRelu3d(2 * b, 2 * b)
        self.conv3 = _ConvBnRelu3d(2 * b, 4 * b, 2)
        self.conv4 = _ConvBnRelu3d(4 * b, 4 * b)
        self.conv5 = _ConvBnRelu3d(4 * b, 8 * b, 2)
        self.conv6 = _ConvBnRelu3d(8 * b, 8 * b)
        self.conv7 = _deconv3d(8 * b, 4 * b)
        self.conv9 = _deconv3d(4 * b, 2 * b)
        self.conv11 = _deconv3d(2 * b, b)
        self.prob = nn.Conv3d(b, 1, 3, stride=1, padding=1, bias=False)

    def forward(self, vol):
        x = vol.permute(0, 2, 1, 3, 4)
        c0 = self.conv0(x)
        c2 = self.conv2(self.conv1(c0))
        c4 = self.conv4(self.conv3(c2))
        x = self.conv6(self.conv5(c4))
        x = c4 + self.conv7(x)
        x = c2 + self.conv9(x)
        x = c0 + self.conv11(x)
        return self.prob(x).squeeze(1)


# ------------------------------------------------------------------ decoders
class _Conv3x3(nn.Module):
    """movedepth/layers.py:537-553 -- reflection pad 1 + 3x3 conv (with bias)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(int(cin), int(cout), 3)

    def forward(self, x):
        return self.conv(F.pad(x, (1, 1, 1, 1), mode="reflect"))


class _ConvBlock(nn.Module):
    """movedepth/layers.py:521-534 -- Conv3x3 + ELU."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _Conv3x3(cin, cout)

    def forward(self, x):
        return F.elu(self.conv(x))


class DepthDecoder(nn.Module):
    """movedepth/networks/depth_decoder.py:10-101 as the trainer instantiates it (trainer.py:74-75:
    skips on, one sigmoid disparity channel per scale).  `decoder` holds the convs in the
    reference's insertion order: for i=4..0 (upconv i 0, upconv i 1), then dispconv per scale."""

    def __init__(self, num_ch_enc, scales=range(4)):
        super().__init__()
        self.scales = list(scales)
        enc = [int(c) for c in num_ch_enc]
        dec = [16, 32, 64, 128, 256]
        mods, self._idx = [], {}
        for i in range(4, -1, -1):
            cin = enc[-1] if i == 4 else dec[i + 1]
            self._idx[("upconv", i, 0)] = len(mods)
            mods.append(_ConvBlock(cin, dec[i]))
            cin = dec[i] + (enc[i - 1] if i > 0 else 0)
            self._idx[("upconv", i, 1)] = len(mods)
            mods.append(_ConvBlock(cin, dec[i]))
        for s in self.scales:
            self._idx[("dispconv", s)] = len(mods)
            mods.append(_Conv3x3(dec[s], 1))
        self.decoder = nn.ModuleList(mods)

    def forward(self, feats):
        out = {}
        x = feats[-1]
        for i in range(4, -1, -1):
            x = self.decoder[self._idx[("upconv", i, 0)]](x)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            if i > 0:
                x = torch.cat([x, feats[i - 1]], 1)
            x = self.decoder[self._idx[("upconv", i, 1)]](x)
            if i in self.scales:
                out[("disp", i)] = torch.sigmoid(self.decoder[self._idx[("dispconv", i)]](x))
        return out


class PoseDecoder(nn.Module):
    """movedepth/networks/pose_decoder.py:8-48 -- squeeze 1x1 -> 2x conv3x3 -> 1x1 -> spatial mean
    -> x0.01 -> (axisangle, translation) each [B,F,1,3]."""

    def __init__(self, num_ch_enc, num_input_features=1, num_frames_to_predict_for=2):
        super().__init__()
        self.nf = num_frames_to_predict_for
        self.net = nn.ModuleList([
            nn.Conv2d(int(num_ch_enc[-1]), 256, 1),
            nn.Conv2d(num_input_features * 256, 256, 3, 1, 1),
            nn.Conv2d(256, 256, 3, 1, 1),
            nn.Conv2d(256, 6 * num_frames_to_predict_for, 1)])

    def forward(self, input_features):
        x = torch.cat([F.relu(self.net[0](f[-1])) for f in input_features], 1)
        x = F.relu(self.net[1](x))
        x = F.relu(self.net[2](x))
        x = self.net[3](x)
        x = 0.01 * x.mean(3).mean(2).view(-1, self.nf, 1, 6)
        return x[..., :3], x[..., 3:]


class UncertNet(nn.Module):
    """movedepth/networks/depth_decoder.py:371-393 -- entropy map -> trust-mono mask.
    The residual `out += x` broadcasts the 1-channel input over 8 channels; written
    out-of-place here (identical values; the in-place form breaks autograd on torch 2.x)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(1, 8, 3, 1, 1, bias=False), nn.BatchNorm2d(8), nn.ReLU())
        self.conv2 = nn.Sequential(nn.Conv2d(8, 8, 3, 1, 1, bias=False), nn.BatchNorm2d(8), nn.ReLU())
        self.head_convs = nn.Conv2d(8, 1, 3, 1, 1, bias=False)

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        out = out + x
        return torch.sigmoid(self.head_convs(out))


class ConvexUpsampleLayer(nn.Module):
    """movedepth/layers.py:184-198 -- mask head conv3x3(no bias)-ReLU-conv1x1(no bias)."""

    def __init__(self, feature_dim, scale=2):
        super().__init__()
        self.scale = scale
        self.upsample_mask = nn.Sequential(
            nn.Conv2d(feature_dim, 64, 3, 1, 1, bias=False), nn.ReLU(),
            nn.Conv2d(64, (2 ** scale) ** 2 * 9, 1, bias=False))

    def forward(self, depth, feat):
        from .layers import convex_upsample
        return convex_upsample(depth, self.upsample_mask(feat), self.scale)
