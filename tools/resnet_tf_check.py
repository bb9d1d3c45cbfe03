"""Teacher-forced per-stage check of the accelerated ResNet forward: every stored tensor of the tape is recomputed by
torch from the tape's OWN inputs of that stage (developer tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodal_ad_b200.models import resnet
from multimodal_ad_b200.models.resnet import _backbone_forward

def nchw(t): return t.float().permute(0, 4, 1, 2, 3)
def rel(a, b): return round(((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item(), 6)
def ulp_ok(a, b): return bool(torch.all((a.float() - b.float()).abs() <= 2 ** -8 * b.float().abs() + 1e-4))

def main(depth=10, n=2, size=32):
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    fn = {10: resnet.resnet10, 18: resnet.resnet18}[depth]
    model = fn(sample_input_D=size, sample_input_H=size, sample_input_W=size, num_seg_classes=1).cuda().train()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
    x = torch.rand(n, 1, size, size, size, device="cuda")
    feats, tape = _backbone_forward(model, x, True, True)
    torch.cuda.synchronize()
    bf = lambda w: w.detach().to(torch.bfloat16).float()
    def bn(t, m): return F.batch_norm(t, None, None, m.weight, m.bias, True, 0.1, m.eps)
    st = tape["stem"]
    c0 = F.conv3d(x.to(torch.bfloat16).float(), bf(model.conv1.weight), stride=2, padding=3)
    print("stem c0", rel(nchw(st["c0"]), c0), ulp_ok(nchw(st["c0"]), c0))
    a0 = F.relu(bn(nchw(st["c0"]), model.bn1))
    print("stem a0", rel(nchw(st["a0"]), a0), ulp_ok(nchw(st["a0"]), a0))
    p0 = F.max_pool3d(nchw(st["a0"]), 3, 2, 1)
    print("stem p0", rel(nchw(tape["blocks"][0]["xin"]), p0))
    for i, r in enumerate(tape["blocks"]):
        b, s, d = r["blk"], r["stride"], r["dil"]
        xin = nchw(r["xin"])
        c1 = F.conv3d(xin, bf(b.conv1.weight), stride=s, padding=d, dilation=d)
        a1 = F.relu(bn(nchw(r["c1"]), b.bn1))
        c2 = F.conv3d(nchw(r["a1"]), bf(b.conv2.weight), padding=d, dilation=d)
        if "cd" in r:
            cd = F.conv3d(xin, bf(b.downsample[0].weight), stride=b.downsample[0].stride)
            res = bn(nchw(r["cd"]), b.downsample[1])
            print(f"blk{i} cd", rel(nchw(r["cd"]), cd), ulp_ok(nchw(r["cd"]), cd))
        else:
            res = xin
        out = F.relu(bn(nchw(r["c2"]), b.bn2) + res)
        print(f"blk{i} c1", rel(nchw(r["c1"]), c1), ulp_ok(nchw(r["c1"]), c1), "a1", rel(nchw(r["a1"]), a1), ulp_ok(nchw(r["a1"]), a1),
              "c2", rel(nchw(r["c2"]), c2), ulp_ok(nchw(r["c2"]), c2), "out", rel(nchw(r["out"]), out), ulp_ok(nchw(r["out"]), out))
    print("feats vs last out", rel(feats, tape["blocks"][-1]["out"]))

if __name__ == "__main__":
    main(*[int(v) for v in sys.argv[1:]])
