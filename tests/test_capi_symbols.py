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


def test_ctypes_signatures_match_the_header(built_lib):
    """Every entry point's ctypes argtypes (multimodal_ad_b200/_lib.py) has as many parameters as its declaration in
    include/mmad_b200.h, with pointers where the header has pointers - a drifted binding would otherwise only show up as a crash
    on the GPU box."""
    import ctypes as C

    from multimodal_ad_b200 import _lib

    lib = _lib.load()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mmad_b200.h")).read(), flags=re.S)
    decls = re.findall(r"\b(?:const\s+char\*|int64_t|int)\s+(mmad_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    assert len(decls) >= 50
    checked = 0
    for name, params in decls:
        plist = [p.strip() for p in params.replace("\n", " ").split(",")]
        if plist == ["void"] or plist == [""]:
            plist = []
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        assert len(fn.argtypes) == len(plist), (name, len(fn.argtypes), plist)
        for at, p in zip(fn.argtypes, plist):
            is_ptr = "*" in p
            at_ptr = at in (C.c_void_p, C.c_char_p) or hasattr(at, "contents") or getattr(at, "_type_", None) is not None and hasattr(at, "_length_")
            if is_ptr:
                assert at_ptr, (name, p, at)
            else:
                assert not at_ptr, (name, p, at)
                want64 = p.startswith("int64_t") or p.startswith("double")
                assert (C.sizeof(at) == 8) == want64, (name, p, at)
        checked += 1
    assert checked >= 45
