"""Host logic of the ROI plan (run programme, work items, slot layout), through the
C-ABI with host_only=1: no CUDA call is made, so this runs on the CPU box."""
import numpy as np
import pytest

from oracle.roi_oracle import roi_pool_oracle, synthetic_atlas
from roi_helpers import emulate_kernel, mean_tolerance


def _plan(lab, r, tile=128, cw=0):
    from multimodal_ad_b200.models.ROI_pol import RoiPlan

    return RoiPlan(lab, r, tile=tile, consumer_warps=cw, host_only=True)


def test_programme_reconstructs_the_label_map(built_lib):
    lab = synthetic_atlas((17, 13, 19), 40, seed=5, empty=(7, 8))
    for tile, cw in ((128, 8), (256, 8), (512, 8), (256, 16)):
        plan = _plan(lab, 40, tile, cw)
        words, offs, ns, smem = plan.programme()
        assert plan.consumer_warps == cw and 2 <= ns <= 4 and smem <= 227 * 1024
        flat = lab.reshape(-1)
        rec = np.zeros_like(flat)
        n_tiles = len(offs) - 1
        assert n_tiles == (flat.size + tile - 1) // tile
        for t in range(n_tiles):
            w0 = offs[t] * 4
            cnt = int(words[w0])
            assert np.all(words[w0 + 1:w0 + 4] == 0)
            assert 0 <= (offs[t + 1] - offs[t]) * 4 - (4 + cnt) < 4
            for run in words[w0 + 4: w0 + 4 + cnt]:
                l, q, ln = int(run >> 24), int((run >> 12) & 0xfff), int(run & 0xfff) + 1
                assert q + ln <= tile and ln <= 8
                assert np.all(rec[t * tile + q: t * tile + q + ln] == 0)
                rec[t * tile + q: t * tile + q + ln] = l
        assert np.array_equal(rec, flat)


@pytest.mark.parametrize("n_vols,sms,cw", [(1, 148, 8), (5, 148, 16), (33, 148, 8), (64, 148, 16), (70, 4, 8),
                                           (200, 3, 16)])
def test_emulated_kernel_matches_oracle(built_lib, n_vols, sms, cw):
    rng = np.random.default_rng(n_vols)
    lab = synthetic_atlas((9, 11, 13), 21, seed=n_vols, empty=(3,))
    feats = rng.standard_normal((n_vols, lab.size)).astype(np.float32)
    plan = _plan(lab, 21, cw=cw)
    mean, mx, arg, cnt = emulate_kernel(plan, feats, sms)
    omean, omx, oarg, ocnt = roi_pool_oracle(feats, lab, 21)
    assert np.array_equal(cnt, ocnt)
    assert np.array_equal(arg, oarg) and np.array_equal(mx, omx)
    assert np.all(np.abs(mean.astype(np.float64) - omean) <= mean_tolerance(feats, lab, 21))


def test_emulated_kernel_random_labels_and_ties(built_lib):
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 256, size=1000).astype(np.int32)        # every run has length ~1, R = 255
    feats = rng.integers(0, 4, size=(3, 1000)).astype(np.float32)  # many ties: first occurrence must win
    plan = _plan(lab, 255)
    mean, mx, arg, cnt = emulate_kernel(plan, feats, 148)
    omean, omx, oarg, ocnt = roi_pool_oracle(feats, lab, 255)
    assert np.array_equal(arg, oarg) and np.array_equal(mx, omx) and np.array_equal(cnt, ocnt)
    assert np.allclose(mean, omean, rtol=1e-6)


def test_binding_covers_every_tile_once(built_lib):
    lab = synthetic_atlas((20, 20, 20), 30, seed=9, empty=())
    plan = _plan(lab, 30)
    n_tiles = (lab.size + 127) // 128
    for n_vols, sms in [(64, 148), (300, 148), (5000, 148), (64, 7)]:
        b = plan.binding(n_vols, sms)
        assert b["grid"] == min(b["n_items"], sms)
        for g in range(b["n_groups"]):
            it = np.flatnonzero(b["item_group"] == g)
            order = np.argsort(b["item_t0"][it])
            t0, t1 = b["item_t0"][it][order], b["item_t1"][it][order]
            assert t0[0] == 0 and t1[-1] == n_tiles and np.array_equal(t0[1:], t1[:-1])
        # every slot is written by exactly one (item, label) and read back by exactly one (group, ROI)
        assert sorted(b["slot_dst"].tolist()) == list(range(b["n_slots"]))
        assert b["fin_ptr"][0] == 0 and b["fin_ptr"][-1] == b["n_slots"] and np.all(np.diff(b["fin_ptr"]) >= 0)


def test_plan_argument_errors(built_lib):
    from multimodal_ad_b200 import _lib
    from multimodal_ad_b200.models.ROI_pol import RoiPlan

    with pytest.raises(_lib.MmadError, match="label outside"):
        RoiPlan(np.array([0, 1, 9], np.int32), 3, host_only=True)
    with pytest.raises(_lib.MmadError, match="n_rois"):
        RoiPlan(np.array([0, 1], np.int32), 300, host_only=True)
    with pytest.raises(_lib.MmadError, match="tile"):
        RoiPlan(np.array([0, 1], np.int32), 1, tile=100, host_only=True)
    with pytest.raises(ValueError):
        RoiPlan(np.zeros(0, np.int32), 1, host_only=True)
