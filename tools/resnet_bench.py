"""Times one ResNet3D-18 training step (forward + CE loss + backward + Adam) on synthetic 1x128^3 volumes (developer tool)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from multimodal_ad_b200.models.Resnet3D import generate_model
from multimodal_ad_b200 import _lib

def main(batch=16, size=128, steps=5, depth=18, warmup=2):
    torch.manual_seed(0)
    model = generate_model(model_depth=depth, input_W=size, input_H=size, input_D=size, nb_class=3, pretrain_path=None,
                           dropout_rate=0.5, device=torch.device("cuda", 0))
    model.train()
    if os.environ.get("MMAD_NO_SIDE"):
        model.wgrad_side_stream = False
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-4, fused=True)
    crit = nn.CrossEntropyLoss()
    x = torch.rand(batch, 1, size, size, size, device="cuda")
    y = torch.randint(0, 3, (batch,), device="cuda")
    def step():
        out = model(x)
        loss = crit(out, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        return loss
    for _ in range(warmup):
        l = step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        l = step()
    enq_ms = (time.perf_counter() - t0) * 1e3 / steps      # host time to enqueue a step (== step time when host-bound)
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / steps
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    # FLOPs of the convolutions (forward), x3 for fwd + dgrad + wgrad
    print(json.dumps(dict(batch=batch, size=size, ms_per_step=round(ms, 3), host_ms=round(host_ms, 3), enqueue_ms=round(enq_ms, 3), vol_per_s=round(batch / ms * 1e3, 1),
                          loss=float(l.detach()), launches_per_step=(_lib.launch_count() - l0) / steps,
                          mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))), flush=True)

if __name__ == "__main__":
    main(*[int(v) for v in sys.argv[1:]])
