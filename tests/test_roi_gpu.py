"""GPU parity of ROI pooling: CUDA path (through the C-ABI) vs the oracle, the
reference fixtures, and size-independent properties at the BASELINE size."""
import os

import numpy as np
import pytest
import torch

from oracle.roi_oracle import roi_mean_backward_oracle, roi_pool_oracle, synthetic_atlas
from roi_helpers import c_oracle_pool, mean_tolerance

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_golden.npz"))
CASES = sorted({k.split("__")[0] for k in GOLD.files})


def _check(mean, mx, arg, feats2d, lab, r, c_oracle=None):
    if c_oracle is not None:
        omean, omx, oarg, ocnt = c_oracle_pool(c_oracle, feats2d, lab, r)
    else:
        omean, omx, oarg, ocnt = roi_pool_oracle(feats2d, lab, r)
    assert np.array_equal(arg, oarg), "argmax indices must be bit-exact"
    assert np.array_equal(mx, omx), "max values must be bit-exact"
    err = np.abs(mean.astype(np.float64) - omean.astype(np.float64))
    tol = mean_tolerance(feats2d, lab, r, rel=1e-6)          # north star: ROI means within 1e-6 relative
    assert np.all(err <= tol), f"mean off by up to {np.max(err / np.maximum(tol, 1e-300)):.3g} x tolerance"
    return ocnt


@pytest.mark.parametrize("case", CASES)
def test_golden_fixture_through_module(case, built_lib):
    from multimodal_ad_b200 import ROIPool

    lab, feats, ref = GOLD[f"{case}__labels"], GOLD[f"{case}__feats"], GOLD[f"{case}__roi_feat"]
    pool = ROIPool(lab)
    x = torch.from_numpy(feats).cuda()
    got = pool(x)                                             # (B,R,C) like image_features.py:114
    assert got.shape == ref.shape
    b, c = feats.shape[:2]
    r = int(lab.max())
    tol = mean_tolerance(feats.reshape(b * c, -1), lab, r).reshape(b, c, r).transpose(0, 2, 1)
    assert np.all(np.abs(got.cpu().numpy().astype(np.float64) - ref) <= tol + 1e-7 * np.abs(ref))
    out = pool.pool(x)
    f2 = feats.reshape(b * c, -1)
    back = lambda t: t.permute(0, 2, 1).reshape(b * c, r).cpu().numpy()   # noqa: E731
    cnt = _check(back(out["mean"]), back(out["max"]), back(out["argmax"]), f2, lab, r)
    assert np.array_equal(out["counts"].numpy(), cnt), "voxel counts must be bit-exact"


@pytest.mark.parametrize("cw", [8, 16])
@pytest.mark.parametrize("tile", [128, 256, 512])
@pytest.mark.parametrize("n_vols", [1, 5, 33, 64, 70])
def test_aal_sized_atlas_vs_c_oracle(n_vols, tile, cw, built_lib, c_oracle):
    from multimodal_ad_b200 import RoiPlan

    if (tile != 256 or cw != 8) and n_vols not in (5, 64):
        pytest.skip("tile / warp sweep on two batch sizes only")
    lab = synthetic_atlas()
    plan = RoiPlan(lab, 170, tile=tile, consumer_warps=cw)
    g = torch.Generator(device="cuda").manual_seed(n_vols)
    x = torch.rand((n_vols, lab.size), device="cuda", generator=g)       # ScaleIntensityd range (ADNI.py:148)
    mean, mx, arg = plan.pool(x)
    torch.cuda.synchronize()
    cnt = _check(mean.cpu().numpy(), mx.cpu().numpy(), arg.cpu().numpy(), x.cpu().numpy(), lab, 170, c_oracle)
    assert np.array_equal(plan.counts, cnt)
    # bit-reproducible
    mean2, mx2, arg2 = plan.pool(x)
    assert torch.equal(mean, mean2) and torch.equal(mx, mx2) and torch.equal(arg, arg2)


@pytest.mark.parametrize("v", [1, 3, 127, 128, 129, 1000, 1001, 1002, 1003, 4099])
def test_ragged_sizes_and_alignments(v, built_lib):
    """Every V mod 4, tiles that end early, and a base pointer that is only 4-byte aligned."""
    from multimodal_ad_b200 import RoiPlan

    rng = np.random.default_rng(v)
    lab = np.repeat(rng.integers(0, 9, size=(v + 10) // 11), 11)[:v].astype(np.int32)   # runs of 11: records of 8 + 3
    plan = RoiPlan(lab, 8, tile=128, consumer_warps=16 if v % 2 else 8)
    for n, off in [(3, 0), (37, 1), (4, 3)]:
        buf = torch.randn(n * v + 4, device="cuda")
        x = buf[off:off + n * v].view(n, v)
        mean, mx, arg = plan.pool(x)
        _check(mean.cpu().numpy(), mx.cpu().numpy(), arg.cpu().numpy(), x.cpu().numpy(), lab, 8)


def test_random_labels_255_rois_and_ties(built_lib):
    from multimodal_ad_b200 import RoiPlan

    rng = np.random.default_rng(7)
    lab = rng.integers(0, 256, size=50_001).astype(np.int32)
    x = torch.from_numpy(rng.integers(0, 5, size=(40, lab.size)).astype(np.float32)).cuda()
    plan = RoiPlan(lab, 255)
    mean, mx, arg = plan.pool(x)
    _check(mean.cpu().numpy(), mx.cpu().numpy(), arg.cpu().numpy(), x.cpu().numpy(), lab, 255)


def test_edge_label_maps(built_lib):
    from multimodal_ad_b200 import RoiPlan

    x = torch.randn(2, 700, device="cuda")
    for lab in (np.zeros(700, np.int32), np.full(700, 4, np.int32),
                np.concatenate([np.zeros(699, np.int32), [2]]).astype(np.int32)):
        plan = RoiPlan(lab, 4, tile=128)
        mean, mx, arg = plan.pool(x)
        _check(mean.cpu().numpy(), mx.cpu().numpy(), arg.cpu().numpy(), x.cpu().numpy(), lab, 4)
    m, a, b = RoiPlan(np.ones(8, np.int32), 1).pool(torch.empty(0, 8, device="cuda"))
    assert m.shape == (0, 1) and a.shape == (0, 1) and b.shape == (0, 1)
    x[0, 5] = float("inf")
    x[1, :] = float("-inf")
    lab = np.full(700, 1, np.int32)
    mean, mx, arg = RoiPlan(lab, 1, tile=128).pool(x)
    assert mx[0, 0].item() == float("inf") and arg[0, 0].item() == 5
    assert mx[1, 0].item() == float("-inf") and arg[1, 0].item() == 0


def test_baseline_config_properties(built_lib):
    """BASELINE.json configs[1] at full size (batch 64 of 91x109x91, 170 labels):
    checks that do not need an oracle pass."""
    from multimodal_ad_b200 import RoiPlan

    lab = synthetic_atlas()
    plan = RoiPlan(lab, 170)
    x = torch.rand((64, lab.size), device="cuda")
    mean, mx, arg = plan.pool(x)
    labt = torch.from_numpy(lab.reshape(-1)).cuda().long()
    cnt = torch.from_numpy(plan.counts).cuda()
    assert torch.equal(cnt.long(), torch.bincount(labt, minlength=171)[1:])
    used = cnt > 0
    # the argmax voxel carries the ROI's label and the max value
    a = arg[:, used].long()
    assert torch.equal(labt[a], (torch.nonzero(used).flatten() + 1).expand_as(a))
    assert torch.equal(torch.gather(x, 1, a), mx[:, used])
    assert torch.all(arg[:, ~used] == -1) and torch.all(mean[:, ~used] == 0)
    # checksum of checksums: sum_r mean_r * count_r == sum of all labelled voxels
    tot = (x.double() * (labt > 0)).sum(1)
    assert torch.allclose((mean.double() * cnt).sum(1), tot, rtol=1e-6)
    # linearity: pool(a*x + y) == a*pool(x) + pool(y)
    y = torch.rand_like(x)
    m2, _, _ = plan.pool(0.5 * x + y)
    my, _, _ = plan.pool(y)
    assert torch.allclose(m2, 0.5 * mean + my, rtol=2e-6, atol=1e-7)
    # mean <= max
    assert torch.all(mean[:, used] <= mx[:, used] + 1e-6)


def test_host_buffer_api_matches_device_api(built_lib):
    from multimodal_ad_b200 import RoiPlan

    lab = synthetic_atlas((31, 37, 29), 50, seed=4, empty=(9,))
    plan = RoiPlan(lab, 50)
    xh = torch.rand(21, lab.size).pin_memory()
    mean, mx, arg = plan.pool(xh.cuda())
    hmean, hmx, harg = plan.pool_host(xh)
    assert np.array_equal(hmean, mean.cpu().numpy()) and np.array_equal(hmx, mx.cpu().numpy())
    assert np.array_equal(harg, arg.cpu().numpy())
    hmean2, _, _ = plan.pool_host(xh.numpy().copy())                      # pageable numpy works too
    assert np.array_equal(hmean2, hmean)


def test_autograd_matches_reference_expression(built_lib):
    from multimodal_ad_b200 import ROIPool
    from oracle.roi_oracle import reference_expression_torch

    lab = synthetic_atlas((7, 9, 11), 10, seed=6, empty=(4,))
    x = torch.randn(2, 3, 7, 9, 11)
    xr = x.clone().requires_grad_(True)
    ref = reference_expression_torch(xr, lab)
    g = torch.randn_like(ref)
    ref.backward(g)
    xc = x.cuda().requires_grad_(True)
    out = ROIPool(lab)(xc)
    out.backward(g.cuda())
    assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(xc.grad.cpu(), xr.grad, rtol=1e-6, atol=1e-9)
    r = 10
    want = roi_mean_backward_oracle(g.permute(0, 2, 1).reshape(6, r).numpy(), lab, r)
    assert np.allclose(xc.grad.cpu().numpy().reshape(6, -1), want, rtol=1e-6, atol=1e-12)


def test_uncropped_feature_map_equals_crop_then_pool(built_lib):
    """image_features.py:103-114 crops the padded UNet feature map to the atlas grid before pooling; ROIPool on the UNCROPPED
    map (atlas embedded with background labels) must give the same means / max / argmax / counts and the same gradient."""
    from multimodal_ad_b200 import ROIPool

    lab = synthetic_atlas((13, 10, 7), 12, seed=11, empty=(5,))
    g = torch.Generator().manual_seed(5)
    big = torch.randn((2, 3, 16, 12, 8), generator=g).cuda()          # padded grid (as 96x112x96 pads 91x109x91)
    crop = big[..., :13, :10, :7].contiguous()
    pool = ROIPool(lab)
    a, b = pool.pool(crop), pool.pool(big)
    assert torch.equal(a["argmax"], b["argmax"]) and torch.equal(a["max"], b["max"]) and torch.equal(a["counts"], b["counts"])
    assert torch.allclose(a["mean"], b["mean"], rtol=1e-6, atol=1e-7)
    f2 = crop.reshape(6, -1).cpu().numpy()
    back = lambda t: t.permute(0, 2, 1).reshape(6, 12).cpu().numpy()   # noqa: E731
    _check(back(b["mean"]), back(b["max"]), back(b["argmax"]), f2, lab, 12)
    xb, xc = big.clone().requires_grad_(True), crop.clone().requires_grad_(True)
    w = torch.randn((2, 12, 3), generator=g).cuda()
    (pool(xb) * w).sum().backward()
    (pool(xc) * w).sum().backward()
    assert torch.allclose(xb.grad[..., :13, :10, :7], xc.grad, rtol=1e-6, atol=1e-9)
    rest = xb.grad.clone()
    rest[..., :13, :10, :7] = 0
    assert rest.abs().max().item() == 0.0                              # nothing flows into the padding
    with pytest.raises(ValueError):
        pool(big[..., :12, :, :])                                      # smaller than the atlas
