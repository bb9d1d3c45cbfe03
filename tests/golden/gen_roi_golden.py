"""Generates tests/golden/roi_golden.npz by executing the reference's own ROI
expressions (image_features.py:80-82,111-114, restated verbatim in
oracle.roi_oracle.reference_expression_torch because the script itself is not
importable) with torch on CPU.  Run from the repo root:

    python tests/golden/gen_roi_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.roi_oracle import reference_expression_torch, synthetic_atlas  # noqa: E402


def main():
    torch.manual_seed(1234)
    rng = np.random.default_rng(1234)
    cases = {}

    def add(name, labels, feats):
        ref = reference_expression_torch(feats, labels).numpy()
        cases[f"{name}__labels"] = labels.astype(np.int32)
        cases[f"{name}__feats"] = feats.numpy()
        cases[f"{name}__roi_feat"] = ref.astype(np.float32)

    # AAL-like small atlas with an unused label (like AAL3's 35/36/81/82), MRI-like [0,1] intensities
    add("shell", synthetic_atlas((13, 11, 9), 12, seed=1, empty=(5,)), torch.rand(2, 3, 13, 11, 9))
    # odd everything, signed features (UNet feature maps are signed), voxel count == 1 (mod 4)
    add("odd", rng.integers(0, 7, size=(7, 5, 3)).astype(np.int32), torch.randn(3, 2, 7, 5, 3))
    # dense parcellation: no background at all, random labels (worst case run structure)
    lab = rng.integers(1, 20, size=(6, 6, 6)).astype(np.int32)
    add("dense", lab, torch.randn(1, 4, 6, 6, 6))
    # long runs, max label with a single voxel
    lab = np.zeros((4, 8, 32), np.int32)
    lab[1:3, 2:6, :] = 3
    lab[3, 7, 31] = 9
    add("runs", lab, torch.rand(5, 1, 4, 8, 32))
    out = os.path.join(os.path.dirname(__file__), "roi_golden.npz")
    np.savez_compressed(out, **cases)
    print("wrote", out, os.path.getsize(out), "bytes;", len(cases) // 3, "cases")


if __name__ == "__main__":
    main()
