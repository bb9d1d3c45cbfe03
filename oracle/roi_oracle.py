"""CPU oracle for atlas ROI pooling.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path
(multimodal_ad_b200/) never does; it fails loudly without the CUDA library.

What it restates
----------------
The reference computes atlas ROI features inline in /root/reference/
image_features.py (models/ROI_pol.py is an empty file in the reference):

  image_features.py:67-69   atlas -> int label volume, roi ids = labels > 0
  image_features.py:80-82   onehot = F.one_hot(labels, max_label+1)[..., 1:]
                            -> (R, D, H, W) float mask, R = max label
  image_features.py:111-112 num = (feats[:,None] * onehot[None,:,None]).sum(spatial)
  image_features.py:113     den = onehot.sum(spatial).clamp_min(1e-6)
  image_features.py:114     roi_feat = num / den            # (B, R, C)

`reference_expression_torch` below re-executes those tensor expressions
verbatim (the script itself is not importable: it runs at import time, reads
absolute paths and needs monai/nibabel).  `roi_pool_oracle` is the restated
algorithm: a segmented reduction over the label map in float64, which is the
exact value the fp32 reference approximates.

Parity status: PINNED against `reference_expression_torch` run in this
container (tests/golden/gen_roi_golden.py writes the fixtures; tests/
test_roi_oracle.py re-checks them), means within 1e-6 relative, counts exact.

Extension beyond the reference (north_star asks for mean/max pooling): the
per-ROI max and argmax.  The reference has no code for them, so they are
defined here: max over the ROI's member voxels, argmax = flat C-order voxel
index (d*H*W + h*W + w) of the FIRST maximal voxel; an empty ROI (count 0)
gives mean 0 (the reference's 0/1e-6), max 0 and argmax -1.
"""
from __future__ import annotations

import numpy as np


def roi_counts(labels: np.ndarray, n_rois: int) -> np.ndarray:
    """Voxel count of every label 1..n_rois (image_features.py:113 `den`
    before the clamp).  int64[n_rois]."""
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    if lab.size and (lab.min() < 0 or lab.max() > n_rois):
        raise ValueError("label outside [0, n_rois]")
    return np.bincount(lab, minlength=n_rois + 1)[1:n_rois + 1]


def roi_pool_oracle(feats: np.ndarray, labels: np.ndarray, n_rois: int):
    """Segmented mean / max / argmax of `feats` (N, V) float32 over `labels`
    (V,) ints in [0, n_rois]; label 0 is background.

    Returns mean float32 (N, R), max float32 (N, R), argmax int32 (N, R),
    counts int64 (R,).  Sums are taken in float64 (exact to ~1e-16), the
    division is the reference's fp32 num / clamp_min(den, 1e-6)
    (image_features.py:113-114).
    """
    f = np.ascontiguousarray(feats, dtype=np.float32)
    if f.ndim != 2:
        raise ValueError("feats must be (N, V)")
    n, v = f.shape
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    if lab.size != v:
        raise ValueError("labels size != V")
    counts = roi_counts(lab, n_rois)
    order = np.argsort(lab, kind="stable")          # voxels grouped by label, ascending index inside
    sorted_lab = lab[order]
    starts = np.searchsorted(sorted_lab, np.arange(1, n_rois + 1), side="left")
    ends = np.searchsorted(sorted_lab, np.arange(1, n_rois + 1), side="right")
    mean = np.zeros((n, n_rois), np.float32)
    mx = np.zeros((n, n_rois), np.float32)
    arg = np.full((n, n_rois), -1, np.int32)
    den = np.maximum(counts.astype(np.float32), np.float32(1e-6))
    for r in range(n_rois):
        idx = order[starts[r]:ends[r]]
        if idx.size == 0:
            continue
        seg = f[:, idx]                                   # (N, cnt), ascending voxel index
        s = seg.astype(np.float64).sum(axis=1)
        mean[:, r] = s.astype(np.float32) / den[r]
        a = np.argmax(seg, axis=1)                        # first occurrence
        mx[:, r] = seg[np.arange(n), a]
        arg[:, r] = idx[a].astype(np.int32)
    return mean, mx, arg, counts


def roi_mean_backward_oracle(grad_mean: np.ndarray, labels: np.ndarray, n_rois: int) -> np.ndarray:
    """d(mean)/d(feats): grad (N, R) -> (N, V); voxel v of ROI r receives
    grad[:, r] / clamp_min(count_r, 1e-6), background receives 0 (autograd of
    image_features.py:111-114)."""
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    counts = roi_counts(lab, n_rois)
    den = np.maximum(counts.astype(np.float32), np.float32(1e-6))
    g = np.asarray(grad_mean, np.float32) / den[None, :]
    g0 = np.concatenate([np.zeros((g.shape[0], 1), np.float32), g], axis=1)
    return g0[:, lab]


def reference_onehot_torch(labels3d):
    """image_features.py:67-69,80-82, verbatim: the (R, D, H, W) float one-hot
    mask the reference builds ONCE, outside its batch loop."""
    import torch
    import torch.nn.functional as F

    aal_data = np.asarray(labels3d).astype(int)
    roi_ids = np.unique(aal_data)
    roi_ids = roi_ids[roi_ids > 0]
    onehot = F.one_hot(torch.from_numpy(aal_data).long(),
                       num_classes=int(roi_ids.max()) + 1)[..., 1:]
    onehot = onehot.permute(3, 0, 1, 2).float()         # (R,D,H,W)
    return onehot


def reference_pool_torch(feats64, onehot):
    """image_features.py:111-114, verbatim: the per-batch ROI pooling.
    Materialises a (B, R, C, D, H, W) product exactly like the reference."""
    num = (feats64[:, None, :, :, :, :] *
           onehot[None, :, None, :, :, :]).sum((-1, -2, -3))
    den = onehot[None, :, None, :, :, :].sum((-1, -2, -3)).clamp_min(1e-6)
    roi_feat = (num / den)                              # (B,R,C)
    return roi_feat


def reference_expression_torch(feats5d, labels3d):
    """The reference's own tensor expressions end to end.  feats5d
    (B, C, D, H, W) float32 torch tensor, labels3d (D, H, W) integer numpy
    array.  Returns roi_feat (B, R, C) with R = labels.max().  Keep
    B*R*C*D*H*W small.
    """
    return reference_pool_torch(feats5d, reference_onehot_torch(labels3d))


def synthetic_atlas(shape=(91, 109, 91), n_rois=170, seed=0, empty=(35, 36, 81, 82)):
    """AAL3-like synthetic label volume: an ellipsoidal grey-matter shell
    partitioned into `n_rois` Voronoi cells, background 0 elsewhere.  Like
    AAL3 (max label 170, labels 35/36/81/82 unused), the labels in `empty`
    have no voxels.  int32 (D, H, W).
    """
    rng = np.random.default_rng(seed)
    d, h, w = shape
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, d), np.linspace(-1, 1, h),
                             np.linspace(-1, 1, w), indexing="ij")
    rad = np.sqrt((zz / 0.80) ** 2 + (yy / 0.84) ** 2 + (xx / 0.78) ** 2)
    shell = (rad < 1.0) & (rad > 0.62)                 # cortex-like shell, ~24 % of the grid
    used = [r for r in range(1, n_rois + 1) if r not in set(empty)]
    # seeds on the shell
    pts = rng.normal(size=(len(used), 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    pts *= rng.uniform(0.66, 0.96, size=(len(used), 1))
    pts *= np.array([0.80, 0.84, 0.78])
    coords = np.stack([zz[shell], yy[shell], xx[shell]], axis=1).astype(np.float32)
    lab = np.zeros(shape, np.int32)
    best = np.full(coords.shape[0], np.inf, np.float32)
    which = np.zeros(coords.shape[0], np.int32)
    for i, p in enumerate(pts.astype(np.float32)):
        dist = ((coords - p) ** 2).sum(axis=1)
        m = dist < best
        best[m] = dist[m]
        which[m] = used[i]
    lab[shell] = which
    return lab
