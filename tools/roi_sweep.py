"""Times mmad_roi_pool_f32 over (tile, stages) on the BASELINE config; prints one line per config.
GPU-only developer tool (not part of the product path)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_ad_b200 import RoiPlan
from oracle.roi_oracle import synthetic_atlas

def main():
    lab = synthetic_atlas()
    V = lab.size
    batches = [int(a) for a in sys.argv[1:]] or [64]
    for B in batches:
        bufs = [torch.rand((B, V), device="cuda") for _ in range(3)]
        for tile, stages, cw in [(t, st, c) for c in (8, 16) for t in (128, 256, 512) for st in (2, 3, 4)]:
            if True:
                try:
                    plan = RoiPlan(lab, 170, tile=tile, stages=stages, consumer_warps=cw)
                except Exception as e:
                    print("skip", tile, stages, e); continue
                _, _, ns, smem = plan.programme()
                if ns != stages: continue
                for i in range(5): plan.pool(bufs[i % 3])
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                K = 100
                a.record()
                for i in range(K): plan.pool(bufs[i % 3])
                b.record(); b.synchronize()
                us = a.elapsed_time(b) * 1e3 / K
                gbs = plan.algorithmic_bytes(B) / (us * 1e-6) / 1e9
                print(json.dumps(dict(batch=B, tile=tile, cw=cw, stages=ns, smem=smem, us=round(us, 2), gbs=round(gbs, 1))), flush=True)
                del plan

if __name__ == "__main__":
    main()
