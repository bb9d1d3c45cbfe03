"""Runs one large wgrad / fwd conv a few times (for ncu captures); prints CUDA-event timings and TFLOP/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ad_b200.models.resnet import _Run

def main(which="wgrad", N=16, S=16, Cin=512, Cout=512, k=3, stride=1, dil=4, reps=5):
    r = _Run(torch.device("cuda", 0))
    pad = dil if k == 3 else 0
    x = torch.randn((N, S, S, S, Cin), device="cuda").to(torch.bfloat16)
    So = (S + 2 * pad - dil * (k - 1) - 1) // stride + 1
    dy = torch.randn((N, So, So, So, Cout), device="cuda").to(torch.bfloat16)
    w = torch.randn((Cout, k ** 3, Cin), device="cuda").to(torch.bfloat16)
    gw = torch.empty((Cout, Cin, k, k, k), device="cuda")
    flops = 2.0 * N * So ** 3 * Cout * Cin * k ** 3
    def f():
        if which == "wgrad": r.wgrad(x, dy, Cout, k, stride, pad, dil, gw)
        else: r.conv(x, w, Cout, k, stride, pad, dil, which == "fwdstats")
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps(dict(which=which, cfg=[N, S, Cin, Cout, k, stride, dil], ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1))), flush=True)

if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], *[int(v) for v in a[1:]])
