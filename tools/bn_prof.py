"""Times the BatchNorm backward kernels (reduce + finalize + apply through _Run.bn_bwd) at the layer shapes of the
ResNet3D-18 bench; prints CUDA-event time and effective GB/s of the two big passes together (developer tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ad_b200.models.resnet import _Run, _p

def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps

def main():
    r = _Run(torch.device("cuda", 0))
    for c, rows in ((64, 524288), (128, 65536), (256, 65536), (512, 65536), (64, 4194304)):
        # several copies so that consecutive launches do not hit the L2 (126 MB)
        ncopy = 1 if rows > 1 << 21 else 4
        xs = [torch.randn((rows, c), device="cuda").to(torch.bfloat16) for _ in range(ncopy * 4)]
        vec = torch.rand((4, c), device="cuda") + 0.5
        it = [0]
        for name, kw in (("dy+mask", dict(use_dy2=False, use_mask=True, mx=False)), ("dy+dy2+mask", dict(use_dy2=True, use_mask=True, mx=False)),
                         ("dy,mask_from_x", dict(use_dy2=False, use_mask=False, mx=True))):
            def f():
                k = (it[0] % ncopy) * 4; it[0] += 1
                r.bn_bwd(xs[k], xs[k + 1] if kw["use_dy2"] else None, xs[k + 2] if kw["use_mask"] else None, xs[k + 3], vec, vec[0], True,
                         mask_from_x=kw["mx"])
            ms = timeit(f)
            nin = 2 + kw["use_dy2"] + kw["use_mask"]
            bytes_ = rows * c * 2 * (nin + 1 + 3)       # reduce: nin reads + g write; apply: g, x reads + dx write
            print(json.dumps(dict(C=c, rows=rows, case=name, ms=round(ms, 4), gbs=round(bytes_ / ms / 1e6, 1))), flush=True)

if __name__ == "__main__":
    main()
