"""Shared helpers for the ROI tests (test infrastructure; may use oracle/)."""
import numpy as np


def mean_tolerance(feats2d, labels, n_rois, rel=1e-6):
    """|mean - ref| bound: `rel` x the ROI's mean |x| (relative 1e-6 as the north
    star states it, made well defined for signed data) + a 1e-30 floor."""
    lab = np.asarray(labels).reshape(-1)
    a = np.abs(np.asarray(feats2d, np.float64))
    tol = np.zeros((a.shape[0], n_rois))
    for r in range(1, n_rois + 1):
        m = lab == r
        if m.any():
            tol[:, r - 1] = a[:, m].mean(axis=1)
    return rel * tol + 1e-30


def c_oracle_pool(lib, feats2d, labels, n_rois):
    f = np.ascontiguousarray(feats2d, np.float32)
    lab = np.ascontiguousarray(np.asarray(labels).reshape(-1), np.int32)
    n, v = f.shape
    mean = np.empty((n, n_rois), np.float32)
    mx = np.empty((n, n_rois), np.float32)
    arg = np.empty((n, n_rois), np.int32)
    cnt = np.empty(n_rois, np.int64)
    rc = lib.roi_pool_oracle_c(f.ctypes.data, n, v, lab.ctypes.data, n_rois, mean.ctypes.data, mx.ctypes.data,
                               arg.ctypes.data, cnt.ctypes.data)
    assert rc == 0
    return mean, mx, arg, cnt


def emulate_kernel(plan, feats2d, sms=148):
    """Executes the plan's run programme + work-item binding exactly the way
    csrc/roi_pool.cu does (same partition of labels over consumer warps, same
    slot layout, same finalize order) in numpy.  Validates the host logic on
    CPU; arithmetic is float64 so it is compared with the oracle, not bit-wise
    with the GPU."""
    words, offs, _, _ = plan.programme()
    n, V = feats2d.shape
    R, T, NW = plan.n_rois, plan.tile, plan.consumer_warps
    HDR = (NW + 1 + 3) // 4 * 4
    b = plan.binding(n, sms)
    ssum = np.zeros((b["n_slots"], 32))
    smax = np.full((b["n_slots"], 32), -np.inf, np.float32)
    sarg = np.full((b["n_slots"], 32), -1, np.int64)
    counts = np.zeros(R, np.int64)
    for item in range(b["n_items"]):
        g = b["item_group"][item]
        vols = np.arange(g * 32, min(n, g * 32 + 32))
        bins_s = np.zeros((R, 32)); bins_m = np.full((R, 32), -np.inf, np.float32); bins_a = np.full((R, 32), -1, np.int64)
        for t in range(b["item_t0"][item], b["item_t1"][item]):
            w0 = offs[t] * 4
            hdr = words[w0:w0 + HDR]
            runs = words[w0 + HDR: offs[t + 1] * 4]
            for w in range(NW):
                prev = -1
                for run in runs[hdr[w]:hdr[w + 1]]:
                    label, q, ln = int(run >> 24), int((run >> 12) & 0xfff), int(run & 0xfff) + 1
                    assert label % NW == w and 1 <= label <= R and 1 <= ln <= 8
                    assert int(run) > prev                      # sorted by (label, start)
                    prev = int(run)
                    seg = feats2d[vols, t * T + q: t * T + q + ln]
                    bins_s[label - 1, :len(vols)] += seg.astype(np.float64).sum(1)
                    for j in range(ln):
                        v = seg[:, j]
                        upd = (v > bins_m[label - 1, :len(vols)]) | (bins_a[label - 1, :len(vols)] < 0)
                        bins_m[label - 1, :len(vols)][upd] = v[upd]
                        bins_a[label - 1, :len(vols)][upd] = t * T + q + j
                    if g == 0:
                        counts[label - 1] += ln
        for j in range(b["item_slot_ptr"][item], b["item_slot_ptr"][item + 1]):
            l = b["slot_label"][j]
            ssum[j], smax[j], sarg[j] = bins_s[l - 1], bins_m[l - 1], bins_a[l - 1]
        touched = set(b["slot_label"][b["item_slot_ptr"][item]:b["item_slot_ptr"][item + 1]].tolist())
        for l in range(1, R + 1):                      # every label the item saw must own a slot
            if bins_a[l - 1, 0] >= 0:
                assert l in touched
    mean = np.zeros((n, R), np.float32); mx = np.zeros((n, R), np.float32); arg = np.full((n, R), -1, np.int32)
    den = np.maximum(counts.astype(np.float32), np.float32(1e-6))
    for g in range(b["n_groups"]):
        for r in range(R):
            s = np.zeros(32); m = np.full(32, -np.inf, np.float32); a = np.full(32, -1, np.int64)
            for k in range(b["fin_ptr"][g * R + r], b["fin_ptr"][g * R + r + 1]):
                sl = b["fin_slots"][k]
                s += ssum[sl]
                upd = (sarg[sl] >= 0) & ((a < 0) | (smax[sl] > m))
                m[upd] = smax[sl][upd]; a[upd] = sarg[sl][upd]
            for lane in range(32):
                vol = g * 32 + lane
                if vol < n:
                    mean[vol, r] = np.float32(s[lane]) / den[r]
                    mx[vol, r] = m[lane] if counts[r] else 0.0
                    arg[vol, r] = a[lane] if counts[r] else -1
    return mean, mx, arg, counts
