"""Generates tests/golden/unet3d_golden.npz from the LIVE reference (/root/reference/models/unet3d.py): default-initialised
UNet3D(1, 1) under a fixed seed, one small input, outputs / hooked tensor / a few gradients in train and eval mode.
Run in the build container (the reference is not present on the GPU box):  python tests/golden/gen_unet_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from test_unet_oracle import _small_target, load_reference  # noqa: E402

SEED, TARGET = 1234, (16, 16, 16)


def main():
    mod = load_reference()
    _small_target(mod, TARGET)
    torch.manual_seed(SEED)
    net = mod.UNet3D(in_channels=1, num_classes=1)
    out = {"seed": np.int64(SEED), "target": np.array(TARGET, np.int64)}
    # the 19 M parameters are not stored: the drop-in's constructor reproduces them from the seed (same module order, same
    # default initialisers), pinned here by one checksum per tensor
    out["keys"] = np.array(list(net.state_dict().keys()))
    out["checksums"] = np.array([float(v.detach().double().abs().sum()) for v in net.state_dict().values()])
    out["shapes"] = np.array([";".join(map(str, v.shape)) for v in net.state_dict().values()])
    g = torch.Generator().manual_seed(7)
    x = torch.rand((2, 1, 13, 15, 11), generator=g)
    wgt = torch.randn((2, 1, 13, 15, 11), generator=g)
    out["x"], out["wgt"] = x.numpy(), wgt.numpy()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    grabbed = {}
    net.s_block1.conv2.register_forward_hook(lambda m, i, o: grabbed.__setitem__("x", o.detach()))
    for training in (True, False):
        tag = "train" if training else "eval"
        net.load_state_dict(sd0)
        net.train(training)
        net.zero_grad(set_to_none=True)
        y = net(x)
        (y * wgt).sum().backward()
        out[f"{tag}.out"] = y.detach().numpy().copy()
        out[f"{tag}.hook"] = grabbed["x"][:, ::8].numpy().copy()          # every 8th channel of the hooked tensor
        named = dict(net.named_parameters())
        for k in ("a_block1.conv1.weight", "s_block1.bn.weight", "s_block1.upconv1.weight", "s_block1.conv3.bias"):
            out[f"{tag}.grad.{k}"] = named[k].grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "unet3d_golden.npz"), **out)
    print("wrote unet3d_golden.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB uncompressed")


if __name__ == "__main__":
    main()
