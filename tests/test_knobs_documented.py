"""Every MMAD_* environment knob the library or the host code reads is listed in INTEGRATION.md (section 5)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts), encoding="utf-8") as f:
        return f.read()


def test_every_environment_knob_is_documented():
    knobs = set()
    csrc = os.path.join(ROOT, "multimodal_ad_b200", "csrc")
    for name in os.listdir(csrc):
        if name.endswith((".cu", ".cuh")):
            knobs |= set(re.findall(r'getenv\("(MMAD_[A-Z0-9_]+)"\)', _read("multimodal_ad_b200", "csrc", name)))
    for rel in (("multimodal_ad_b200", "_lib.py"), ("multimodal_ad_b200", "sharding.py"), ("multimodal_ad_b200", "models", "resnet.py"),
                ("multimodal_ad_b200", "models", "unet3d.py"), ("bench.py",)):
        knobs |= set(re.findall(r'environ(?:\.get)?[\(\[]\s*"(MMAD_[A-Z0-9_]+)"', _read(*rel)))
    assert knobs, "no knobs found: the patterns above no longer match the sources"
    doc = _read("INTEGRATION.md")
    missing = sorted(k for k in knobs if k not in doc)
    assert not missing, f"undocumented environment knobs: {missing}"
