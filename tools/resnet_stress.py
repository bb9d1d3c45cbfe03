"""Crash / sanity sweep of the accelerated backbone over depths, shortcut types, batch sizes and volume shapes: one training step
each, outputs and gradients must be finite (developer tool)."""
import os, sys, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ad_b200.models import resnet

def main():
    torch.manual_seed(0)
    cases = [(10, "B", 1, (64, 64, 64)), (18, "B", 3, (40, 50, 60)), (18, "A", 5, (91, 109, 91)), (34, "B", 2, (96, 112, 96)),
             (50, "B", 3, (80, 98, 80)), (18, "B", 7, (33, 47, 29)), (101, "B", 1, (48, 48, 48)), (18, "B", 32, (64, 64, 64)),
             (18, "B", 1, (160, 192, 160)), (34, "A", 4, (17, 19, 23))]
    ok = True
    for depth, sc, n, shape in cases:
        fn = {10: resnet.resnet10, 18: resnet.resnet18, 34: resnet.resnet34, 50: resnet.resnet50, 101: resnet.resnet101}[depth]
        m = fn(sample_input_D=shape[0], sample_input_H=shape[1], sample_input_W=shape[2], num_seg_classes=1, shortcut_type=sc).cuda()
        m.train()
        x = torch.rand((n, 1) + shape, device="cuda")
        f = m.features(x)
        f.square().mean().backward()
        torch.cuda.synchronize()
        fin = bool(torch.isfinite(f).all()) and all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.backbone_parameters())
        gn = float(sum(p.grad.double().square().sum() for p in m.backbone_parameters()).sqrt())
        print(json.dumps(dict(depth=depth, shortcut=sc, batch=n, shape=shape, feats=list(f.shape), finite=fin, grad_norm=round(gn, 4),
                              mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))), flush=True)
        ok &= fin and gn > 0
        del m, x, f
        torch.cuda.empty_cache()
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)

if __name__ == "__main__":
    main()
