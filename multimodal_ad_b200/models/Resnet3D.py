"""Model factory — drop-in for /root/reference/models/Resnet3D.py:6-113 and train_ResNet3D.py:44-84.

`generate_model` keeps the reference's arguments.  The segmentation head is replaced by global average pooling +
Linear exactly as the reference does (Resnet3D.py:85-86 / train_ResNet3D.py:66-71); pretrained weights are loaded when
the file exists (train_ResNet3D.py:75-83).  nn.DataParallel wrapping (Resnet3D.py:89-99) is not reproduced: multi-GPU
runs use one process per GPU (bench.py / torch.distributed)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import resnet

_FC_IN = {10: 256, 18: 512, 34: 512, 50: 2048, 101: 2048, 152: 2048, 200: 2048}


def generate_model(model_type='resnet', model_depth=18, input_W=112, input_H=136, input_D=112, resnet_shortcut='B',
                   no_cuda=False, gpu_id=(0,), pretrain_path='config/pretrain/resnet_18_23dataset.pth', nb_class=2,
                   dropout_rate=None, device=None):
    assert model_type in ['resnet']
    assert model_depth in [10, 18, 34, 50, 101, 152, 200]
    fn = {10: resnet.resnet10, 18: resnet.resnet18, 34: resnet.resnet34, 50: resnet.resnet50, 101: resnet.resnet101,
          152: resnet.resnet152, 200: resnet.resnet200}[model_depth]
    model = fn(sample_input_W=input_W, sample_input_H=input_H, sample_input_D=input_D, shortcut_type=resnet_shortcut,
               no_cuda=no_cuda, num_seg_classes=1)
    # resnet10's layer4 has 512 channels like resnet18 (the reference's 256 would not match its own model); use the real width
    fc_in = 512 * model.block_type.expansion
    head = [nn.AdaptiveAvgPool3d((1, 1, 1)), nn.Flatten()]
    if dropout_rate is not None:
        head.append(nn.Dropout(p=dropout_rate))                    # train_ResNet3D.py:69
    head.append(nn.Linear(in_features=fc_in, out_features=nb_class, bias=True))
    model.conv_seg = nn.Sequential(*head)
    if device is None:
        device = torch.device('cpu') if no_cuda else torch.device('cuda', gpu_id[0])
    model.to(device)
    if pretrain_path and os.path.isfile(pretrain_path):
        ckpt = torch.load(pretrain_path, map_location=device)
        state = ckpt.get('state_dict', ckpt)
        sd = model.state_dict()
        sd.update({k.replace('module.', ''): v for k, v in state.items() if k.replace('module.', '') in sd})
        model.load_state_dict(sd)
    return model
