"""The C-ABI library loads on a CPU-only box and exports every symbol
include/*.h declares (no compute call is made here)."""
import ctypes
import glob
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(mmad_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    declared = _declared()
    assert len(declared) >= 10
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/ but not exported: {missing}"
    lib.mmad_abi_version.restype = ctypes.c_int
    assert lib.mmad_abi_version() == 2
    lib.mmad_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mmad_last_error(), bytes)       # the LAST failure of this thread: earlier tests may have provoked one


def test_product_path_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "multimodal_ad_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
    for path in glob.glob(os.path.join(ROOT, "multimodal_ad_b200", "csrc", "*.cu*")):
        assert "oracle/" not in open(path).read().replace("// oracle/", ""), path


def test_cpu_tensor_raises_instead_of_falling_back(built_lib):
    import numpy as np
    import pytest
    import torch

    from multimodal_ad_b200 import _lib
    from multimodal_ad_b200.models.ROI_pol import ROIPool

    pool = ROIPool(np.ones((2, 2, 2), np.int32))
    with pytest.raises(_lib.MmadError, match="CUDA"):
        pool(torch.zeros(1, 1, 2, 2, 2))
