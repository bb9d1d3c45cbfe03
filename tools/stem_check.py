"""GPU check of the space-to-depth stem (mmad_stem_s2d_*) against torch conv3d on the same bf16 operands (developer tool)."""
import os, sys, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodal_ad_b200 import _lib
from multimodal_ad_b200.models.resnet import _Run, _p

def run(N, D, H, W, seed=0, reps=0):
    r = _Run(torch.device("cuda", 0)); lib = r.lib
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((N, 1, D, H, W), device="cuda", generator=g)
    w = torch.randn((64, 1, 7, 7, 7), device="cuda", generator=g) / 343 ** 0.5
    Do, Ho, Wo = (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    xs = r.empty((lib.mmad_stem_s2d_elems(N, D, H, W),))
    wk = r.empty((64, 512))
    y = r.empty((N, Do, Ho, Wo, 64))
    npart = lib.mmad_stem_s2d_stats_partials(N, D, H, W)
    part = r.empty((npart, 64, 2), torch.float32)
    r.chk(lib.mmad_stem_s2d_pack(_p(x), _p(xs), N, D, H, W, r.stream), "pack")
    r.chk(lib.mmad_stem_s2d_prep_weights(_p(w), _p(wk), r.stream), "prep")
    r.chk(lib.mmad_stem_s2d_fwd(_p(xs), _p(wk), _p(y), _p(part), N, D, H, W, r.stream), "fwd")
    torch.cuda.synchronize()
    xr = x.to(torch.bfloat16).float().requires_grad_(False)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv3d(xr, wr, stride=2, padding=3)
    out = dict(cfg=[N, D, H, W])
    yf = y.float().permute(0, 4, 1, 2, 3)
    out["fwd_max"] = (yf - ref).abs().max().item()
    out["fwd_ref"] = ref.abs().mean().item()
    st = part.double().sum(0)
    out["stat_sum"] = (st[:, 0] - yf.double().sum((0, 2, 3, 4))).abs().max().item() / max(1.0, yf.double().sum((0, 2, 3, 4)).abs().max().item())
    out["stat_sq"] = ((st[:, 1] - (yf.double() ** 2).sum((0, 2, 3, 4))).abs() / (yf.double() ** 2).sum((0, 2, 3, 4))).max().item()
    # wgrad
    dy = torch.randn((N, Do, Ho, Wo, 64), device="cuda", generator=g).to(torch.bfloat16)
    ns = ctypes.c_int(0)
    elems = lib.mmad_stem_s2d_wgrad_workspace(N, D, H, W, ctypes.byref(ns))
    ws = r.empty((elems,), torch.float32)
    gw = torch.empty_like(w)
    r.chk(lib.mmad_stem_s2d_wgrad(_p(xs), _p(dy), _p(ws), N, D, H, W, r.stream), "wgrad")
    r.chk(lib.mmad_stem_s2d_wgrad_reduce(_p(ws), ns.value, _p(gw), r.stream), "reduce")
    torch.cuda.synchronize()
    ref.backward(dy.float().permute(0, 4, 1, 2, 3))
    out["wgrad_rel"] = ((gw - wr.grad).norm() / wr.grad.norm()).item()
    out["ok"] = bool(out["fwd_max"] < 0.02 * max(1.0, out["fwd_ref"] * 4) and out["wgrad_rel"] < 2e-3 and out["stat_sum"] < 1e-4 and out["stat_sq"] < 1e-4)
    if reps:
        for name, f in (("fwd", lambda: lib.mmad_stem_s2d_fwd(_p(xs), _p(wk), _p(y), _p(part), N, D, H, W, r.stream)),
                        ("wgrad", lambda: lib.mmad_stem_s2d_wgrad(_p(xs), _p(dy), _p(ws), N, D, H, W, r.stream)),
                        ("pack", lambda: lib.mmad_stem_s2d_pack(_p(x), _p(xs), N, D, H, W, r.stream))):
            f(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps): f()
            b.record(); b.synchronize()
            out[name + "_ms"] = round(a.elapsed_time(b) / reps, 4)
    print(json.dumps(out), flush=True)
    return out["ok"]

if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfgs = [(1, 16, 16, 16), (2, 8, 8, 16), (1, 13, 11, 19), (2, 32, 32, 32), (1, 91, 109, 91)]
    a = [int(v) for v in sys.argv[1:]]
    ok = True
    if len(a) >= 4:
        ok = run(*a[:4], reps=a[4] if len(a) > 4 else 0)
    else:
        for c in cfgs: ok &= run(*c)
    print("ALL OK" if ok else "FAILURES")
