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


def _pack_key(v, idx):
    v = np.float32(v) + np.float32(0.0)
    u = int(np.frombuffer(np.float32(v).tobytes(), np.uint32)[0])
    u = (~u & 0xffffffff) if (u & 0x80000000) else (u | 0x80000000)
    return (u << 32) | (0xffffffff - int(idx))


def _unpack_key(key):
    u = key >> 32
    u = (u & 0x7fffffff) if (u & 0x80000000) else (~u & 0xffffffff)
    return np.frombuffer(np.uint32(u).tobytes(), np.float32)[0], 0xffffffff - (key & 0xffffffff)


def emulate_kernel(plan, feats2d, sms=148):
    """Executes the plan's run programme + work-item binding exactly the way
    csrc/roi_pool.cu does (records split evenly over the consumer warps, per-warp
    register partial folded into the CTA accumulators on ROI change, packed
    max/argmax keys, slot layout, finalize order) in numpy.  Validates the host
    logic on CPU; sums are float64 so it is compared with the oracle, not
    bit-wise with the GPU."""
    words, offs, _, _ = plan.programme()
    n, V = feats2d.shape
    R, T, NW, HDR = plan.n_rois, plan.tile, plan.consumer_warps, 4
    b = plan.binding(n, sms)
    ssum = np.zeros((b["n_slots"], 32))
    skey = np.zeros((b["n_slots"], 32), dtype=object)
    written = np.zeros(b["n_slots"], bool)
    counts = np.zeros(R, np.int64)
    for item in range(b["n_items"]):
        g = b["item_group"][item]
        vols = np.arange(g * 32, min(n, g * 32 + 32))
        bins_s = np.zeros((R, 32)); bins_k = np.zeros((R, 32), dtype=object)
        state = [dict(cur=0, ds=np.zeros(32), mx=np.full(32, -np.inf, np.float32), arg=np.full(32, -1, np.int64))
                 for _ in range(NW)]

        def flush(st):
            if st["cur"]:
                for ln in range(len(vols)):
                    if st["arg"][ln] >= 0:
                        bins_s[st["cur"] - 1, ln] += st["ds"][ln]
                        bins_k[st["cur"] - 1, ln] = max(bins_k[st["cur"] - 1, ln], _pack_key(st["mx"][ln], st["arg"][ln]))

        for t in range(b["item_t0"][item], b["item_t1"][item]):
            w0 = offs[t] * 4
            cnt = int(words[w0])
            recs = words[w0 + HDR: w0 + HDR + cnt]
            assert np.all(np.diff(recs.astype(np.int64)) > 0)           # sorted by (label, start)
            assert (offs[t + 1] - offs[t]) * 4 - (HDR + cnt) in (0, 1, 2, 3)
            for w in range(NW):
                st = state[w]
                for rec in recs[cnt * w // NW: cnt * (w + 1) // NW]:
                    label, q, ln = int(rec >> 24), int((rec >> 12) & 0xfff), int(rec & 0xfff) + 1
                    assert 1 <= label <= R and 1 <= ln <= 8 and q + ln <= T
                    if label != st["cur"]:
                        flush(st)
                        st.update(cur=label, ds=np.zeros(32), mx=np.full(32, -np.inf, np.float32),
                                  arg=np.full(32, -1, np.int64))
                    seg = feats2d[vols, t * T + q: t * T + q + ln]
                    st["ds"][:len(vols)] += seg.astype(np.float64).sum(1)
                    a = seg.argmax(1)                                     # first occurrence inside the record
                    v = seg[np.arange(len(vols)), a]
                    upd = (v > st["mx"][:len(vols)]) | (st["arg"][:len(vols)] < 0)
                    st["mx"][:len(vols)][upd] = v[upd]
                    st["arg"][:len(vols)][upd] = (t * T + q + a)[upd]
                    if g == 0:
                        counts[label - 1] += ln
        for st in state:
            flush(st)
        lo, hi = b["item_slot_ptr"][item], b["item_slot_ptr"][item + 1]
        touched = set(b["slot_label"][lo:hi].tolist())
        for j in range(lo, hi):
            l, d = b["slot_label"][j], b["slot_dst"][j]
            assert not written[d]
            written[d] = True
            ssum[d], skey[d] = bins_s[l - 1], bins_k[l - 1]
            assert b["fin_ptr"][g * R + l - 1] <= d < b["fin_ptr"][g * R + l]
        for l in range(1, R + 1):                      # every label the item saw must own a slot
            if any(k != 0 for k in bins_k[l - 1]):
                assert l in touched
    assert written.all()
    mean = np.zeros((n, R), np.float32); mx = np.zeros((n, R), np.float32); arg = np.full((n, R), -1, np.int32)
    den = np.maximum(counts.astype(np.float32), np.float32(1e-6))
    for g in range(b["n_groups"]):
        for r in range(R):
            s = np.zeros(32); key = [0] * 32
            for k in range(b["fin_ptr"][g * R + r], b["fin_ptr"][g * R + r + 1]):
                s += ssum[k]
                key = [max(a_, b_) for a_, b_ in zip(key, skey[k])]
            for lane in range(32):
                vol = g * 32 + lane
                if vol < n:
                    mean[vol, r] = np.float32(s[lane]) / den[r]
                    if counts[r] and key[lane]:
                        mx[vol, r], arg[vol, r] = _unpack_key(key[lane])
    return mean, mx, arg, counts
