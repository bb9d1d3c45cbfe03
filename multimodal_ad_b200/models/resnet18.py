"""Drop-in for the reference's models/resnet18.py (3D-ResNet-18 backbone with dilated layer3/4 and a segmentation head)
on the B200 kernels: same class / factory names, constructor arguments, module construction order, initialisation calls
and state_dict keys as /root/reference/models/resnet18.py:74-184.  conv1 ... layer4 (resnet18.py:168-173) run inside the
accelerated autograd Function of models/resnet.py; `conv_seg` (resnet18.py:112-120) stays a torch module."""
from __future__ import annotations

import torch.nn as nn

from . import resnet as _accel
from .resnet import BasicBlock, conv3x3x3, downsample_basic_block   # noqa: F401  (re-exported like the reference)

__all__ = ["ResNet", "resnet18"]


class ResNet(_accel.ResNet):
    # resnet18.py:74-175
    def __init__(self, block, layers, sample_input_D, sample_input_H, sample_input_W, num_seg_classes=1, shortcut_type="B",
                 no_cuda=False):
        nn.Module.__init__(self)                       # the parent's constructor builds resnet.py's variant of the head
        self.inplanes = 64
        self.no_cuda = no_cuda
        self.block_type = block
        self.shortcut_type = shortcut_type
        self.conv1 = nn.Conv3d(1, 64, kernel_size=7, stride=(2, 2, 2), padding=3, bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0], shortcut_type)
        self.layer2 = self._make_layer(block, 128, layers[1], shortcut_type, stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], shortcut_type, stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], shortcut_type, stride=1, dilation=4)
        self.conv_seg = nn.Sequential(
            nn.ConvTranspose3d(512, 32, 2, stride=2, bias=False),
            nn.BatchNorm3d(32),
            nn.ReLU(inplace=True),
            nn.Conv3d(32, 32, 3, padding=1, bias=False),
            nn.BatchNorm3d(32),
            nn.ReLU(inplace=True),
            nn.Conv3d(32, num_seg_classes, 1, bias=False),
        )
        self._init_weights()

    def _init_weights(self):
        # resnet18.py:159-165
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def resnet18(**kwargs):
    # resnet18.py:178-187
    return ResNet(BasicBlock, [2, 2, 2, 2], **kwargs)
