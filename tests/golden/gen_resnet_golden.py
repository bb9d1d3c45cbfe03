"""Generates tests/golden/resnet10_golden.npz by importing the reference's own models/resnet.py (only possible where
/root/reference is mounted) and running resnet10 on a seeded 16^3 input on CPU in fp32:

    python tests/golden/gen_resnet_golden.py

Stored: input, backbone features (resnet.py:205-212), loss and a few parameter-gradient checksums of
sum(features * weight).  Weights come from the reference's own initialisation under torch.manual_seed(1234), which
this repo's models/resnet.py reproduces (same module construction order, same init calls)."""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))


def load_reference_resnet():
    spec = importlib.util.spec_from_file_location("ref_resnet", "/root/reference/models/resnet.py")
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_resnet()
    torch.manual_seed(1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ref.resnet10(sample_input_D=16, sample_input_H=16, sample_input_W=16, num_seg_classes=1, no_cuda=True)
    g = torch.Generator().manual_seed(99)
    x = torch.rand(2, 1, 16, 16, 16, generator=g)
    wgt = torch.randn(2, 512, 2, 2, 2, generator=g)
    m.train()
    feats = m.layer4(m.layer3(m.layer2(m.layer1(m.maxpool(m.relu(m.bn1(m.conv1(x))))))))
    loss = (feats * wgt).sum()
    loss.backward()
    out = {"x": x.numpy(), "wgt": wgt.numpy(), "features": feats.detach().numpy(), "loss": np.float64(loss.item())}
    for k in ("conv1.weight", "bn1.weight", "layer1.0.conv2.weight", "layer2.0.downsample.0.weight", "layer4.0.bn2.bias"):
        out["grad__" + k] = dict(m.named_parameters())[k].grad.numpy()
    out["init_checksum"] = np.float64(sum(float(p.double().abs().sum()) for p in m.parameters()))
    path = os.path.join(os.path.dirname(__file__), "resnet10_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

    # Bottleneck network (resnet.py:72-109) and shortcut type 'A' (resnet.py:26-37): features and init checksums only
    for name, kw, layers in (("resnet50", dict(), "resnet50"), ("resnet18a", dict(shortcut_type="A"), "resnet18")):
        torch.manual_seed(4321)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = getattr(ref, layers)(sample_input_D=16, sample_input_H=16, sample_input_W=16, num_seg_classes=1, no_cuda=True, **kw)
        init_cs = np.float64(sum(float(p.detach().double().abs().sum()) for p in m.parameters()))
        g = torch.Generator().manual_seed(77)
        x = torch.rand(2, 1, 16, 20, 12, generator=g)
        m.train()
        with torch.no_grad():
            feats = m.layer4(m.layer3(m.layer2(m.layer1(m.maxpool(m.relu(m.bn1(m.conv1(x))))))))
        path = os.path.join(os.path.dirname(__file__), name + "_golden.npz")
        np.savez_compressed(path, x=x.numpy(), features=feats.numpy(), init_checksum=init_cs)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
