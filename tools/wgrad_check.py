"""GPU check of mmad_conv3d_wgrad_bf16 (+ reduce) and dgrad-through-fwd against torch autograd (developer tool)."""
import os, sys, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodal_ad_b200 import _lib
from multimodal_ad_b200.models.resnet import _Run, _p

def run(N, D, H, W, Cin, Cout, k, stride, pad, dil, seed=0):
    r = _Run(torch.device("cuda", 0))
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, k, k, k), device="cuda", generator=g) / (k ** 1.5 * Cin ** 0.5))
    Do = (D + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    dy = torch.randn((N, Do, Ho, Wo, Cout), device="cuda", generator=g).to(torch.bfloat16)
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    y = F.conv3d(xr, wr, stride=stride, padding=pad, dilation=dil)
    y.backward(dy.float().permute(0, 4, 1, 2, 3))
    gw = torch.empty_like(w)
    r.wgrad(x, dy, Cout, k, stride, pad, dil, gw)
    taps = k * k * k
    wf = r.empty((Cout, taps, Cin)); wt = r.empty((Cin, taps, Cout))
    r.chk(r.lib.mmad_conv3d_prep_weights(_p(w), _p(wf), _p(wt), Cout, Cin, taps, r.stream), "prep")
    out = dict(cfg=[N, D, H, W, Cin, Cout, k, stride, pad, dil])
    out["wgrad_rel"] = ((gw - wr.grad).norm() / wr.grad.norm()).item()
    out["wgrad_max"] = (gw - wr.grad).abs().max().item() / wr.grad.abs().max().item()
    if Cin == 384:
        out["dgrad_rel"] = 0.0
        dx = None
    elif stride == 1:
        dx, _ = r.conv(dy, wt, Cin, k, 1, dil * (k - 1) - pad, dil, False)
    else:
        up = r.empty((N, D, H, W, Cout))
        r.chk(r.lib.mmad_upsample_zero2(_p(dy), _p(up), N, Do, Ho, Wo, D, H, W, Cout, r.stream), "up")
        dx, _ = r.conv(up, wt, Cin, k, 1, dil * (k - 1) - pad, dil, False)
    torch.cuda.synchronize()
    ref_dx = xr.grad.permute(0, 2, 3, 4, 1)
    if dx is not None:
        out["dgrad_rel"] = ((dx.float() - ref_dx).norm() / ref_dx.norm()).item()
    out["ok"] = bool(out["wgrad_rel"] < 2e-3 and out["dgrad_rel"] < 1e-2)
    print(json.dumps(out), flush=True)
    return out["ok"]

if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    cfgs = [
        (1, 4, 4, 4, 64, 64, 1, 1, 0, 1),        # one chunk, mode 2 (Cin 64), single tap
        (1, 4, 8, 8, 128, 128, 1, 1, 0, 1),      # mode 1
        (2, 8, 8, 8, 64, 64, 3, 1, 1, 1),        # 27 taps in pairs (mode 2), padding
        (1, 8, 8, 16, 128, 256, 3, 1, 2, 2),     # dilation 2, NB 256
        (1, 16, 16, 16, 64, 128, 3, 2, 1, 1),    # stride 2
        (1, 16, 16, 16, 64, 128, 1, 2, 0, 1),    # 1x1x1 stride 2
        (1, 5, 7, 9, 256, 512, 3, 1, 4, 4),      # ragged, dilation 4, two co tiles
        (1, 1, 1, 4096, 384, 64, 1, 1, 0, 1),    # stem view: rows x 384
        (1, 5, 7, 9, 64, 64, 3, 1, 1, 1),        # halo kernel, ragged
        (3, 23, 28, 23, 64, 64, 3, 1, 1, 1),     # halo kernel, layer1 shape of a 91x109x91 input
        (2, 16, 16, 16, 128, 128, 3, 1, 1, 1),   # pair kernel, odd unit count (phantom unit), batch-deep chunks
        (2, 16, 16, 16, 256, 512, 3, 1, 4, 4),   # pair kernel, dilation 4 with padding skips, two co tiles
        (3, 8, 8, 8, 128, 256, 3, 1, 1, 1),      # pair kernel, odd batch
        (1, 16, 16, 16, 128, 256, 1, 1, 0, 1),   # pair kernel, 1x1x1 (one unit: phantom peer)
    ]
    sel = [int(a) for a in sys.argv[1:]] or range(len(cfgs))
    ok = True
    for i in sel:
        ok &= run(*cfgs[i])
    print("ALL OK" if ok else "FAILURES")
