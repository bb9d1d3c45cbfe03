"""GPU check of mmad_conv3d_fwd_bf16 against torch conv3d (developer tool; tests/test_conv_gpu.py is the suite)."""
import os, sys, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodal_ad_b200 import _lib

def run(N, D, H, W, Cin, Cout, k, stride, pad, dil, seed=0, stats=True):
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((Cout, k * k * k, Cin), device="cuda", generator=g) / (k ** 1.5 * Cin ** 0.5)).to(torch.bfloat16)
    Do = (D + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    y = torch.full((N, Do, Ho, Wo, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    npart = lib.mmad_conv3d_stats_partials(N, D, H, W, Cout, k, stride, pad, dil)
    part = torch.zeros((npart, Cout, 2), device="cuda") if stats else None
    rc = lib.mmad_conv3d_fwd_bf16(x.data_ptr(), w.data_ptr(), y.data_ptr(), part.data_ptr() if stats else None,
                                  N, D, H, W, Cin, Cout, k, stride, pad, dil,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "conv3d_fwd")
    torch.cuda.synchronize()
    xr = x.float().permute(0, 4, 1, 2, 3)
    wr = w.float().reshape(Cout, k, k, k, Cin).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(xr, wr, stride=stride, padding=pad, dilation=dil).permute(0, 2, 3, 4, 1)
    err = (y.float() - ref).abs()
    scale = ref.abs().mean().item()
    out = dict(cfg=[N, D, H, W, Cin, Cout, k, stride, pad, dil], max_err=err.max().item(), mean_err=err.mean().item(),
               ref_mean_abs=scale, nan=int(torch.isnan(y.float()).sum().item()))
    if stats:
        s = part.sum(0)
        yb = y.float().reshape(-1, Cout)
        out["stat_sum_err"] = (s[:, 0] - yb.sum(0)).abs().max().item() / (yb.abs().sum(0).max().item() + 1e-9)
        out["stat_sq_err"] = (s[:, 1] - (yb * yb).sum(0)).abs().max().item() / ((yb * yb).sum(0).max().item() + 1e-9)
    out["ok"] = bool(out["nan"] == 0 and out["max_err"] < 0.03 * max(scale, 1e-3) * 8)
    print(json.dumps(out), flush=True)
    return out["ok"]

if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfgs = [
        (1, 1, 8, 16, 64, 64, 1, 1, 0, 1),      # one tile, 1x1x1: plain GEMM 128x64x64
        (1, 2, 8, 16, 128, 128, 1, 1, 0, 1),    # 2 tiles, K = 2 slices, BN = 128
        (1, 4, 8, 16, 64, 256, 1, 1, 0, 1),     # BN = 256
        (2, 8, 8, 8, 64, 64, 3, 1, 1, 1),       # 3x3x3 pad 1
        (1, 8, 16, 16, 128, 256, 3, 1, 2, 2),   # dilation 2
        (1, 16, 16, 16, 64, 128, 3, 2, 1, 1),   # stride 2
        (1, 16, 16, 16, 64, 128, 1, 2, 0, 1),   # 1x1x1 stride 2 (downsample)
        (1, 5, 7, 9, 64, 64, 3, 1, 1, 1),       # ragged extents (tiles partly out of bounds)
        (1, 6, 11, 23, 64, 512, 3, 1, 4, 4),    # dilation 4, Cout 512 (two N tiles)
        (2, 32, 32, 32, 64, 64, 3, 1, 1, 1),    # many tiles per CTA (persistent loop, TMEM double buffer)
        (1, 1, 1, 4194304, 384, 64, 1, 1, 0, 1),  # the stem GEMM at full size (16 x 64^3 rows x 384)
        (1, 1, 1, 300000, 384, 64, 1, 1, 0, 1),
    ]
    sel = [int(a) for a in sys.argv[1:]] or range(len(cfgs))
    ok = True
    for i in sel:
        ok &= run(*cfgs[i])
    print("ALL OK" if ok else "FAILURES")
