"""Drop-in for the reference's models/ImageEncoder.py (3D-ResNet encoder without a segmentation head, optional global
average pooling) on the B200 kernels: same class / factory names, constructor arguments, construction order,
initialisation and state_dict keys as /root/reference/models/ImageEncoder.py:121-248.  (The reference file itself does
not import: its downsample_basic_block repeats the `device=` keyword, ImageEncoder.py:66-67; the restatement here follows
its text.)  conv1 ... layer4 run inside the accelerated autograd Function of models/resnet.py; BasicBlock encoders
(image_encoder18 / 34) with one input channel are covered, the Bottleneck factories construct but raise on forward."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import resnet as _accel
from .resnet import BasicBlock, Bottleneck, conv3x3x3, downsample_basic_block   # noqa: F401

__all__ = ["ImageEncoder", "image_encoder18", "image_encoder34", "image_encoder50", "image_encoder101", "image_encoder152",
           "image_encoder200"]


class ImageEncoder(_accel.ResNet):
    # ImageEncoder.py:121-221
    def __init__(self, block, layers, in_channels=1, shortcut_type="B", no_cuda=False, global_pool=False):
        nn.Module.__init__(self)
        self.global_pool = global_pool
        self.no_cuda = no_cuda
        self.inplanes = 64
        self.block_type = block
        self.shortcut_type = shortcut_type
        self.in_channels = in_channels
        self.conv1 = nn.Conv3d(in_channels, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0], shortcut_type)
        self.layer2 = self._make_layer(block, 128, layers[1], shortcut_type, stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], shortcut_type, stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], shortcut_type, stride=1, dilation=4)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        # ImageEncoder.py:209-220
        x = self.features(x)
        if self.global_pool:
            x = F.adaptive_avg_pool3d(x, 1)
            x = torch.flatten(x, 1)
        return x


def image_encoder18(**kwargs):
    return ImageEncoder(BasicBlock, [2, 2, 2, 2], **kwargs)


def image_encoder34(**kwargs):
    return ImageEncoder(BasicBlock, [3, 4, 6, 3], **kwargs)


def image_encoder50(**kwargs):
    return ImageEncoder(Bottleneck, [3, 4, 6, 3], **kwargs)


def image_encoder101(**kwargs):
    return ImageEncoder(Bottleneck, [3, 4, 23, 3], **kwargs)


def image_encoder152(**kwargs):
    return ImageEncoder(Bottleneck, [3, 8, 36, 3], **kwargs)


def image_encoder200(**kwargs):
    return ImageEncoder(Bottleneck, [3, 24, 36, 3], **kwargs)
