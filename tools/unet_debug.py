"""Per-stage comparison of the accelerated UNet3D with the oracle (eval folded, eval with autograd, training)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ad_b200.models import unet3d
from multimodal_ad_b200.models.unet3d import tape_stages
from oracle.unet_oracle import unet3d_oracle

def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()

def main():
    target = (32, 48, 32)
    torch.manual_seed(1)
    m = unet3d.UNet3D(1, 1).cuda()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm3d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.3, 0.3)
                mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.5, 1.5)
    m.target = target
    m.keep_tape = True
    x = torch.rand(2, 1, 29, 45, 27, device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    for mode in ("eval_fold", "eval_grad", "train"):
        m.train(mode == "train")
        if mode == "eval_fold":
            with torch.no_grad():
                out = m(x)
        else:
            out = m(x)
        st = {k: v.cpu() for k, v in tape_stages(m, m._last_tape).items()}
        comp = {}
        ref = unet3d_oracle({k: v.clone() for k, v in sd.items()}, x.cpu(), mode == "train", emulate_bf16=True, forced=st, computed=comp, target=target)
        print("==", mode, "out", rel(out.detach().cpu(), ref))
        for k in comp:
            if k in st:
                print(f"   {k:16s} {rel(st[k], comp[k]):.3e}")
        m.load_state_dict(sd)

main()
