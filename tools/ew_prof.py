"""Runs the stem-sized bandwidth kernels a few times (for ncu captures) and prints CUDA-event timings + effective GB/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ad_b200.models.resnet import _Run, _p

def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps

def main():
    r = _Run(torch.device("cuda", 0)); lib = r.lib
    n, s = 16, 128
    x = torch.rand(n, 1, s, s, s, device="cuda")
    so = 64
    rows = n * so ** 3
    a0 = torch.randn((n, so, so, so, 64), device="cuda").to(torch.bfloat16)
    p0 = r.empty((n, 32, 32, 32, 64)); idx = torch.empty((n, 32, 32, 32, 64), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: r.chk(lib.mmad_maxpool3d_fwd(_p(a0), _p(p0), _p(idx), n, so, so, so, 64, r.stream), "mpf"))
    print(json.dumps(dict(k="maxpool_fwd", ms=round(ms, 3), gbs=round((a0.numel() * 2 + p0.numel() * 3) / ms / 1e6, 1))))
    dy = torch.randn_like(p0); dx = r.empty(a0.shape)
    ms = timeit(lambda: r.chk(lib.mmad_maxpool3d_bwd(_p(dy), _p(idx), _p(dx), n, so, so, so, 64, r.stream), "mpb"))
    print(json.dumps(dict(k="maxpool_bwd", ms=round(ms, 3), gbs=round((a0.numel() * 2 + p0.numel() * 3) / ms / 1e6, 1))))
    vec = torch.rand((4, 64), device="cuda")
    ms = timeit(lambda: r.bn_apply(a0, vec, True))
    print(json.dumps(dict(k="bn_apply_stem", ms=round(ms, 3), gbs=round(a0.numel() * 4 / ms / 1e6, 1))))
    ms = timeit(lambda: r.bn_bwd(a0, None, a0, a0, vec, vec[0], True))
    print(json.dumps(dict(k="bn_bwd_stem(reduce+apply)", ms=round(ms, 3), gbs=round(a0.numel() * 2 * 7 / ms / 1e6, 1))))

if __name__ == "__main__":
    main()
