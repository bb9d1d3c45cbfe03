"""Pins the oracle: numpy and C restatements against the fixtures produced by
the reference's own tensor expressions (tests/golden/gen_roi_golden.py)."""
import os

import numpy as np
import pytest

from oracle.roi_oracle import (reference_expression_torch, roi_counts, roi_mean_backward_oracle, roi_pool_oracle,
                               synthetic_atlas)
from roi_helpers import c_oracle_pool, mean_tolerance

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_golden.npz"))
CASES = sorted({k.split("__")[0] for k in GOLD.files})


@pytest.mark.parametrize("case", CASES)
def test_numpy_oracle_matches_reference_fixture(case):
    lab, feats, ref = GOLD[f"{case}__labels"], GOLD[f"{case}__feats"], GOLD[f"{case}__roi_feat"]
    b, c = feats.shape[:2]
    r = int(lab.max())
    mean, mx, arg, cnt = roi_pool_oracle(feats.reshape(b * c, -1), lab, r)
    got = mean.reshape(b, c, r).transpose(0, 2, 1)                       # (B,R,C) like image_features.py:114
    tol = mean_tolerance(feats.reshape(b * c, -1), lab, r).reshape(b, c, r).transpose(0, 2, 1)
    assert got.shape == ref.shape
    assert np.all(np.abs(got.astype(np.float64) - ref) <= tol + 1e-7 * np.abs(ref))
    # counts are what the reference's `den` holds before the clamp
    assert np.array_equal(cnt, [(lab == k).sum() for k in range(1, r + 1)])
    # max / argmax definition
    f2 = feats.reshape(b * c, -1)
    flat = lab.reshape(-1)
    for k in range(1, r + 1):
        idx = np.flatnonzero(flat == k)
        if idx.size == 0:
            assert np.all(mean[:, k - 1] == 0) and np.all(mx[:, k - 1] == 0) and np.all(arg[:, k - 1] == -1)
        else:
            assert np.array_equal(mx[:, k - 1], f2[:, idx].max(1))
            assert np.array_equal(arg[:, k - 1], idx[f2[:, idx].argmax(1)])


@pytest.mark.parametrize("case", CASES)
def test_c_oracle_matches_numpy_oracle(case, c_oracle):
    lab, feats = GOLD[f"{case}__labels"], GOLD[f"{case}__feats"]
    f2 = feats.reshape(feats.shape[0] * feats.shape[1], -1)
    r = int(lab.max())
    a = roi_pool_oracle(f2, lab, r)
    b = c_oracle_pool(c_oracle, f2, lab, r)
    assert np.allclose(a[0], b[0], rtol=2e-7, atol=1e-30)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


def test_reference_expression_reexecuted_live():
    """The fixture generator's expression, re-run here on a fresh input."""
    import torch

    lab = synthetic_atlas((9, 10, 11), 15, seed=3, empty=(2, 15))
    feats = torch.rand(2, 2, 9, 10, 11)
    ref = reference_expression_torch(feats, lab).numpy()
    r = int(lab.max())
    mean, *_ = roi_pool_oracle(feats.reshape(4, -1).numpy(), lab, r)
    assert np.allclose(mean.reshape(2, 2, r).transpose(0, 2, 1), ref, rtol=1e-6, atol=1e-9)


def test_ties_and_first_occurrence(c_oracle):
    lab = np.array([1, 1, 0, 1, 2, 2, 2, 0], np.int32)
    f = np.array([[3, 5, 9, 5, -1, -1, -2, 7]], np.float32)
    for mean, mx, arg, cnt in (roi_pool_oracle(f, lab, 3), c_oracle_pool(c_oracle, f, lab, 3)):
        assert arg.tolist() == [[1, 4, -1]] and mx.tolist() == [[5.0, -1.0, 0.0]]
        assert np.allclose(mean, [[13 / 3, -4 / 3, 0.0]]) and cnt.tolist() == [3, 3, 0]


def test_backward_oracle_is_the_autograd_of_the_reference():
    import torch

    lab = synthetic_atlas((5, 6, 7), 6, seed=2, empty=(4,))
    feats = torch.rand(2, 3, 5, 6, 7, requires_grad=True)
    out = reference_expression_torch(feats, lab)                          # (B,R,C)
    g = torch.randn_like(out)
    out.backward(g)
    r = int(lab.max())
    got = roi_mean_backward_oracle(g.permute(0, 2, 1).reshape(6, r).numpy(), lab, r)
    assert np.allclose(got.reshape(feats.shape), feats.grad.numpy(), rtol=1e-6, atol=1e-9)


def test_label_range_checked():
    with pytest.raises(ValueError):
        roi_counts(np.array([0, 3]), 2)


def test_synthetic_atlas_is_aal3_like():
    lab = synthetic_atlas()
    assert lab.shape == (91, 109, 91) and lab.max() == 170
    assert set(np.unique(lab)) == set(range(171)) - {35, 36, 81, 82}
    assert 0.15 < (lab > 0).mean() < 0.35


REF_SCRIPT = "/root/reference/image_features.py"


@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="reference not mounted on this box")
def test_oracle_is_pinned_to_the_reference_source_text():
    """The script cannot be imported (it runs at import, absolute paths, MONAI), so the pin works on its TEXT: lines 80-82
    (one-hot mask) and 111-114 (masked sum / clamped count) are read from /root/reference/image_features.py, dedented and
    exec'd on a small case; the restatement in oracle/roi_oracle.py and the numpy oracle must reproduce the result."""
    import textwrap

    import torch
    import torch.nn.functional as F

    from oracle.roi_oracle import reference_onehot_torch, reference_pool_torch

    lines = open(REF_SCRIPT, encoding="utf-8").read().split("\n")
    mask_src = textwrap.dedent("\n".join(lines[79:82]))
    pool_src = textwrap.dedent("\n".join(lines[110:114]))
    assert "F.one_hot" in mask_src and "permute(3,0,1,2)" in mask_src, mask_src
    assert "feats64[:,None" in pool_src and "clamp_min(1e-6)" in pool_src and "roi_feat" in pool_src, pool_src
    lab = synthetic_atlas((9, 11, 8), 12, seed=3, empty=(5,))
    g = torch.Generator().manual_seed(11)
    feats = torch.randn((2, 3, 9, 11, 8), generator=g)
    aal_data = np.asarray(lab).astype(int)
    roi_ids = np.unique(aal_data)
    roi_ids = roi_ids[roi_ids > 0]                                       # image_features.py:68
    env = {"F": F, "torch": torch, "aal_data": aal_data, "roi_ids": roi_ids}
    exec(mask_src, env)                                                  # noqa: S102 - the reference's own lines
    env["feats64"] = feats
    exec(pool_src, env)                                                  # noqa: S102
    want = env["roi_feat"]
    assert torch.equal(env["onehot"], reference_onehot_torch(lab))
    assert torch.equal(want, reference_pool_torch(feats, reference_onehot_torch(lab)))
    r = int(lab.max())
    mean, _, _, _ = roi_pool_oracle(feats.numpy().reshape(6, -1), lab, r)
    got = mean.reshape(2, 3, r).transpose(0, 2, 1)
    tol = mean_tolerance(feats.numpy().reshape(6, -1), lab, r).reshape(2, 3, r).transpose(0, 2, 1)
    assert np.all(np.abs(got.astype(np.float64) - want.numpy()) <= tol + 1e-7 * np.abs(want.numpy()))
