"""The `models` shim (multimodal_ad_b200/shim) lets the reference's training script run unchanged: its own
`from models import resnet` (train_ResNet3D.py:19) and `generate_model` (train_ResNet3D.py:44-84), taken from the script's
TEXT, build the accelerated network with the reference's state_dict keys and initialisation."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = "/root/reference"
SCRIPT = os.path.join(REF, "train_ResNet3D.py")

DRIVER = r'''
import sys, textwrap, torch, torch.nn as nn, os
lines = open(sys.argv[1], encoding="utf-8").read().split("\n")
assert lines[18].strip() == "from models import resnet", lines[18]
env = {"torch": torch, "nn": nn, "os": os}
exec(lines[18], env)                                         # train_ResNet3D.py:19, verbatim
exec(textwrap.dedent("\n".join(lines[43:84])), env)          # train_ResNet3D.py:44-84: generate_model, verbatim
torch.manual_seed(7)
net = env["generate_model"](model_depth=18, input_W=91, input_H=109, input_D=91, pretrain_path="/nonexistent", nb_class=2)
print("MODULE", type(net).__module__)
print("KEYS", ",".join(net.state_dict().keys()))
print("CHECKSUM", sum(float(p.detach().double().abs().sum()) for p in net.parameters()))
import models
print("OTHER", getattr(__import__("models.network", fromlist=["x"]), "__file__", "?"))
try:
    import types
    sys.modules.setdefault("torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))   # the reference's unet3d.py imports it
    from models.unet3d import UNet3D                             # image_features.py:6, verbatim
    print("UNET", UNet3D.__module__)
except Exception as e:
    print("UNET", "error:" + type(e).__name__)
'''


def _run(shim, cwd, tmp_path):
    """Runs DRIVER from a file inside `cwd`'s namespace: plainly (the reference's own models/) or through the launcher."""
    drv = tmp_path / ("driver_shim.py" if shim else "driver_ref.py")
    drv.write_text(DRIVER)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, REF]))
    cmd = [sys.executable] + (["-m", "multimodal_ad_b200.run"] if shim else []) + [str(drv), SCRIPT]
    out = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return dict(l.split(" ", 1) for l in out.stdout.strip().split("\n") if " " in l and l.split(" ", 1)[0].isupper())


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="reference not mounted on this box")
def test_reference_script_builds_the_accelerated_network_through_the_shim(tmp_path):
    ref = _run(False, REF, tmp_path)                                             # the reference's own models/resnet.py
    got = _run(True, REF, tmp_path)                                              # same text through python -m multimodal_ad_b200.run
    assert ref["MODULE"] == "models.resnet"
    assert got["MODULE"] == "multimodal_ad_b200.models.resnet"
    assert got["KEYS"] == ref["KEYS"]
    assert abs(float(got["CHECKSUM"]) - float(ref["CHECKSUM"])) <= 1e-9 * float(ref["CHECKSUM"])
    # modules the accelerated path does not provide still resolve to the reference's own files
    assert got["OTHER"].startswith(REF)
    # image_features.py:6 `from models.unet3d import UNet3D`
    assert ref["UNET"] == "models.unet3d" and got["UNET"] == "multimodal_ad_b200.models.unet3d"
