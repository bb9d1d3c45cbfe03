"""Runs a few UNet3D eval forwards (+ ROI pooling) / training steps for ncu launch lists; prints CUDA-event timings."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ad_b200 import RoiPlan
from multimodal_ad_b200.models import unet3d
from oracle.roi_oracle import synthetic_atlas

def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "eval"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    torch.manual_seed(0)
    m = unet3d.UNet3D(1, 1).cuda()
    x = torch.rand(batch, 1, 91, 109, 91, device="cuda")
    plan = RoiPlan(synthetic_atlas((91, 109, 91), 170), 170)
    if mode == "eval":
        m.eval()
        def step():
            with torch.no_grad():
                return m.roi_features(x, plan)
    else:
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-5)
        def step():
            out = m(x)
            loss = (out * out).mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
    for _ in range(2): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): step()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps(dict(mode=mode, batch=batch, ms_per_step=round(ms, 3), volumes_per_s=round(batch / ms * 1e3, 1), mem_gb=round(torch.cuda.max_memory_allocated() / 1e9, 2))))

main()
