"""Encoders and the cost-volume regulariser.

Checkpoint key layout (the compatibility contract) is kept:
`encoder.*` for the ResNet trunks, `conv{0..3}.{i}.{conv,bn}.*` / `inner1.*` / `out.weight` for
FPN4, `conv{0..6}.{conv,bn}.*`, `conv{7,9,11}.{0,1}.*`, `prob.weight` for reg3d.
The dense convolutions run on cuDNN (library GEMMs); what is specific to this build is how
they are fed: `reg3d.forward_volume` takes the fused cost-volume kernel's [B,G,D,h,w] output
directly, so the reference's permute copy (resnet_encoder.py:257) never happens.
"""
import glob
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.models as tvm

from .. import norm as NM
from .. import precision as PR


def _resnet(num_layers, num_input_images):
    ctor = {18: tvm.resnet18, 34: tvm.resnet34, 50: tvm.resnet50, 101: tvm.resnet101, 152: tvm.resnet152}
    if num_layers not in ctor:
        raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
    if num_input_images > 1 and num_layers not in (18, 50):
        raise ValueError("multi-image encoders exist for 18 or 50 layers")
    net = ctor[num_layers](weights=None)
    if num_input_images > 1:
        # pose encoder: conv1 widened to 3*n inputs; every conv re-drawn kaiming(fan_out) as the
        # reference's ResNetMultiImageInput does (resnet_encoder.py:21-44)
        net.conv1 = nn.Conv2d(3 * num_input_images, 64, kernel_size=7, stride=2, padding=3, bias=False)
        for m in net.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
    return net


def _load_imagenet(net, num_layers, num_input_images):
    """ImageNet weights from <repo>/pretrain_resnet/resnet{N}-*.pth if present (the reference
    looks in the same place, resnet_encoder.py:95-102); conv1 is tiled / averaged for n images."""
    root = os.path.join(os.path.dirname(__file__), "..", "..", "pretrain_resnet")
    hits = glob.glob(os.path.join(root, "resnet{}-*.pth".format(num_layers)))
    if not hits:
        raise FileNotFoundError("pretrained=True but no resnet{}-*.pth under {}".format(num_layers, root))
    sd = torch.load(hits[0], map_location="cpu")
    if num_input_images > 1:
        sd["conv1.weight"] = torch.cat([sd["conv1.weight"]] * num_input_images, 1) / num_input_images
    net.load_state_dict(sd)


class ResnetEncoder(nn.Module):
    """ResNet trunk returning 5 feature maps.  Reference: movedepth/networks/resnet_encoder.py:74-121."""

    def __init__(self, num_layers, pretrained, num_input_images=1, **kwargs):
        super().__init__()
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        net = _resnet(num_layers, num_input_images)
        if pretrained:
            _load_imagenet(net, num_layers, num_input_images)
        del net.fc
        del net.avgpool
        self.encoder = NM.adopt(net)             # residual blocks: conv -> fused BN(+ReLU, +residual) kernels
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4

    def forward(self, input_image):
        e = self.encoder
        x = (input_image - 0.45) / 0.225
        f0 = NM.bn_act(e.bn1, e.conv1(x), relu=True)
        mp = e.maxpool
        if (f0.is_cuda and f0.dtype == torch.float32 and f0.shape[1] % 4 == 0 and mp.kernel_size == 3 and mp.stride == 2 and mp.padding == 1
                and mp.dilation == 1 and not mp.ceil_mode):
            from .. import ops
            pooled = ops.maxpool3x3s2(f0)              # hand-written channels-last kernels (ATen: 130 + 200 us per encoder)
        else:
            pooled = mp(f0)
        f1 = e.layer1(pooled)
        f2 = e.layer2(f1)
        f3 = e.layer3(f2)
        f4 = e.layer4(f3)
        self.features = [f0, f1, f2, f3, f4]
        return self.features


class Conv2d(nn.Module):
    """conv (no bias) + BN + ReLU.  Reference: movedepth/networks/resnet_encoder.py:453-475."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, relu=True, bn=True, bn_momentum=0.1, **kw):
        super().__init__()
        self.conv = PR.Conv2d(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kw)
        self.bn = nn.BatchNorm2d(out_channels, momentum=bn_momentum) if bn else None
        self.relu = relu

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            return NM.bn_act(self.bn, x, relu=self.relu)
        return F.relu(x, inplace=True) if self.relu else x


class FPN4(nn.Module):
    """Matching-feature pyramid: returns (matching feature at 1/2**scale, the stage feature there).
    Reference: movedepth/networks/resnet_encoder.py:311-391 (dcn=True needs an external
    deformable-conv extension and is not supported)."""

    def __init__(self, base_channels, scale=0, dcn=False):
        super().__init__()
        if dcn:
            raise NotImplementedError("--dcn needs the external DeformConvPack extension (out of scope)")
        b = base_channels
        self.base_channels, self.scale, self.dcn = b, scale, dcn
        self.conv0 = nn.Sequential(Conv2d(3, b, 3, 1, padding=1), Conv2d(b, b, 3, 1, padding=1))
        stages = []
        for i in range(1, 4):
            cin, cout = b * 2 ** (i - 1), b * 2 ** i
            stages.append(nn.Sequential(Conv2d(cin, cout, 5, stride=2, padding=2), Conv2d(cout, cout, 3, 1, padding=1),
                                        Conv2d(cout, cout, 3, 1, padding=1)))
        self.conv1, self.conv2, self.conv3 = stages
        top = 8 * b
        if scale < 3:
            self.inner1 = PR.Conv2d(4 * b, top, 1, bias=True)
        if scale < 2:
            self.inner2 = PR.Conv2d(2 * b, top, 1, bias=True)
        if scale < 1:
            self.inner3 = PR.Conv2d(b, top, 1, bias=True)
        if scale == 3:
            self.out = PR.Conv2d(top, 8 * b, 1, bias=False)
        else:
            self.out = PR.Conv2d(top, b * 2 ** scale, 3, padding=1, bias=False)

    def forward(self, x):
        feats = [self.conv0(x)]
        for stage in (self.conv1, self.conv2, self.conv3):
            feats.append(stage(feats[-1]))
        y = feats[3]
        for level, lateral in ((2, "inner1"), (1, "inner2"), (0, "inner3")):
            if self.scale <= level:
                y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True) + getattr(self, lateral)(feats[level])
        return self.out(y), feats[self.scale]


class ConvBnReLU3D(nn.Module):
    """Reference: movedepth/networks/resnet_encoder.py:175-182."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = PR.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)

    def forward(self, x):
        c = self.conv
        if (PR.get_policy() == "3xtf32" and self.bn.training and x.is_cuda and x.dim() == 5 and tuple(c.weight.shape) == (16, 16, 3, 3, 3)
                and c.bias is None and tuple(c.stride) == (1, 1, 1) and tuple(c.padding) == (1, 1, 1) and not PR._policy["split_backward"]):
            # reg3d's full-resolution first layer: the tcgen05 conv's epilogue sums the BatchNorm statistics of its output
            from .. import ops
            y, sums = ops.conv3d_c16_to_16_with_stats(x, c.weight, 3)
            PR.record_conv(c.weight, x, y)
            return NM.bn_act(self.bn, y, relu=True, sums=sums)
        return NM.bn_act(self.bn, self.conv(x), relu=True)


class ConvBnReLUSeq(nn.Sequential):
    """nn.Sequential(conv, BatchNorm, ReLU) -- the reference's container and state-dict keys (`{0,1}.*`) -- whose forward
    runs the fused BN+ReLU kernels."""

    def forward(self, x, skip=None):
        """relu(bn(conv(x))) [+ skip]: the U-Net skip addition (resnet_encoder.py:272-276) rides in the BN-apply kernel."""
        return NM.bn_act(self[1], self[0](x), relu=True, residual=skip, post=True)


def _up3d(cin, cout, k=3, p=1, op=1, s=2):
    return ConvBnReLUSeq(PR.ConvTranspose3d(cin, cout, kernel_size=k, padding=p, output_padding=op, stride=s, bias=False),
                         nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


class ProbConv3d(PR.Conv3d):
    """reg3d's output head.  Conv3d(16 -> 1, 3x3x3, pad 1, no bias) on a CUDA volume runs the hand-written exact-fp32
    stencil kernels (ops.conv3d_c16_to_1); any other shape follows the precision policy like every other conv.
    Same parameters / state-dict keys as nn.Conv3d."""

    def forward(self, x):
        if (x.is_cuda and self.in_channels == 16 and self.out_channels == 1 and self.bias is None
                and tuple(self.kernel_size) == (3, 3, 3) and tuple(self.stride) == (1, 1, 1)
                and tuple(self.padding) == (1, 1, 1) and tuple(self.dilation) == (1, 1, 1) and self.groups == 1):
            from .. import ops
            return ops.conv3d_c16_to_1(x, self.weight)
        return super().forward(x)


class _UNet3D(nn.Module):
    def _unet(self, x):
        c0 = self.conv0(x)
        c2 = self.conv2(self.conv1(c0))
        c4 = self.conv4(self.conv3(c2))
        y = self.conv6(self.conv5(c4))
        y = self.conv7(y, c4)
        y = self.conv9(y, c2)
        y = self.conv11(y, c0)
        return self.prob(y).squeeze(1)

    def forward_volume(self, vol_bgdhw):
        """[B,G,D,h,w] (the fused cost-volume kernel's layout) -> logits [B,D,h,w]."""
        return self._unet(vol_bgdhw)

    def forward(self, x):
        """Reference calling convention: [B,D,G,h,w] -> logits [B,D,h,w]."""
        return self._unet(x.permute(0, 2, 1, 3, 4))


class reg3d(_UNet3D):
    """3-level 3-D U-Net regulariser.  Reference: movedepth/networks/resnet_encoder.py:227-280."""

    def __init__(self, in_channels, base_channels, down_size=3):
        super().__init__()
        if down_size != 3:
            raise NotImplementedError("the trainer builds reg3d with down_size=3 only")
        self.down_size = down_size
        b = base_channels
        self.conv0 = ConvBnReLU3D(in_channels, b)
        self.conv1 = ConvBnReLU3D(b, 2 * b, stride=2)
        self.conv2 = ConvBnReLU3D(2 * b, 2 * b)
        self.conv3 = ConvBnReLU3D(2 * b, 4 * b, stride=2)
        self.conv4 = ConvBnReLU3D(4 * b, 4 * b)
        self.conv5 = ConvBnReLU3D(4 * b, 8 * b, stride=2)
        self.conv6 = ConvBnReLU3D(8 * b, 8 * b)
        self.conv7 = _up3d(8 * b, 4 * b)
        self.conv9 = _up3d(4 * b, 2 * b)
        self.conv11 = _up3d(2 * b, b)
        self.prob = ProbConv3d(b, 1, 3, stride=1, padding=1, bias=False)


class reg2d(_UNet3D):
    """Per-depth-slice (1x3x3) variant used when D < 8.  Reference: resnet_encoder.py:184-225."""

    def __init__(self, input_channel=128, base_channel=32):
        super().__init__()
        b = base_channel
        k, p, s = (1, 3, 3), (0, 1, 1), (1, 2, 2)
        self.conv0 = ConvBnReLU3D(input_channel, b, kernel_size=k, pad=p)
        self.conv1 = ConvBnReLU3D(b, 2 * b, kernel_size=k, stride=s, pad=p)
        self.conv2 = ConvBnReLU3D(2 * b, 2 * b)
        self.conv3 = ConvBnReLU3D(2 * b, 4 * b, kernel_size=k, stride=s, pad=p)
        self.conv4 = ConvBnReLU3D(4 * b, 4 * b)
        self.conv5 = ConvBnReLU3D(4 * b, 8 * b, kernel_size=k, stride=s, pad=p)
        self.conv6 = ConvBnReLU3D(8 * b, 8 * b)
        self.conv7 = _up3d(8 * b, 4 * b, k, p, p, s)
        self.conv9 = _up3d(4 * b, 2 * b, k, p, p, s)
        self.conv11 = _up3d(2 * b, b, k, p, p, s)
        self.prob = PR.Conv3d(8, 1, 1, stride=1, padding=0)
