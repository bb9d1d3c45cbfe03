"""GPU check of the whole accelerated ResNet (forward + backward) against the torch fp32 oracle (developer tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodal_ad_b200.models import resnet
from oracle.resnet_oracle import resnet_features_oracle

def main(depth=10, n=2, size=32, seed=0):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(seed)
    fn = {10: resnet.resnet10, 18: resnet.resnet18}[depth]
    model = fn(sample_input_D=size, sample_input_H=size, sample_input_W=size, num_seg_classes=1).cuda()
    layers = [len(l) for l in (model.layer1, model.layer2, model.layer3, model.layer4)]
    # make BN affine parameters non-trivial
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
    x = torch.rand(n, 1, size, size, size, device="cuda")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.train()
    model.keep_tape = True
    feats = model.features(x)
    from multimodal_ad_b200.models.resnet import tape_stages
    forced = tape_stages(model, model._last_tape)
    wgt = torch.randn_like(feats) / feats.numel() ** 0.5
    loss = (feats * wgt).sum()
    loss.backward()
    torch.cuda.synchronize()
    def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
    named = dict(model.named_parameters())
    out = {"feat_shape": list(feats.shape)}
    for emu in ("forced", True, False):
        leaves = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        ref = resnet_features_oracle(leaves, x, layers, True, emulate_bf16=bool(emu), forced=forced if emu == "forced" else None)
        (ref * wgt).sum().backward()
        tag = emu if emu == "forced" else ("emu" if emu else "fp32")
        out[f"feat_rel_{tag}"] = rel(feats, ref)
        errs = []
        for k, v in leaves.items():
            if v.grad is None or k.startswith("conv_seg"): continue
            g = named[k].grad
            errs.append((k, round(rel(g, v.grad), 5) if g is not None else float("nan")))
        out[f"grad_rel_{tag}"] = errs if emu == "forced" else sorted(errs, key=lambda t: -t[1])[:3]
        out[f"grad_rel_median_{tag}"] = sorted(e for _, e in errs)[len(errs) // 2]
        out[f"grad_rel_max_{tag}"] = max(e for _, e in errs)
    # running stats
    out["running_mean_rel"] = rel(model.bn1.running_mean, 0.9 * sd["bn1.running_mean"] + 0.1 * F.conv3d(x, sd["conv1.weight"], stride=2, padding=3).mean(dim=(0, 2, 3, 4)))
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a)
