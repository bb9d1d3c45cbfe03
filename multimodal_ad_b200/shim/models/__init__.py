"""`models` package shim: lets the reference's scripts run UNCHANGED on the accelerated modules.

The reference's scripts import their networks as `from models import resnet` (train_ResNet3D.py:19),
`from models.unet3d import UNet3D` (image_features.py:6), `from models.Resnet3D import generate_model`, ...  Run the script through
the launcher (multimodal_ad_b200/run.py puts this directory in front of the script's own on sys.path),

    cd /path/to/Multimodal_AD && PYTHONPATH=/path/to/repo python -m multimodal_ad_b200.run train_ResNet3D.py

and `models.<name>` resolves to `multimodal_ad_b200.models.<name>` for every module the accelerated path provides;
every other `models.*` module (mymodel, MSHyper, network, ...) still comes from the reference's own `models/` directory,
which is appended to this package's search path when it is found on sys.path / in the working directory.
"""
import importlib
import os
import sys

_ACCELERATED = ("resnet", "resnet18", "ImageEncoder", "Resnet3D", "ROI_pol", "unet3d")

for _name in _ACCELERATED:
    try:
        _mod = importlib.import_module("multimodal_ad_b200.models." + _name)
    except ModuleNotFoundError as _e:                      # a module this build does not ship: leave it to the reference
        if _e.name != "multimodal_ad_b200.models." + _name:
            raise
        continue
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod

_here = os.path.dirname(os.path.abspath(__file__))
for _base in [os.getcwd()] + list(sys.path):
    _cand = os.path.join(_base or os.getcwd(), "models")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)                             # the reference's own models/ for everything not accelerated
