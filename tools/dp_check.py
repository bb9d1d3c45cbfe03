"""2-GPU check of the data-parallel path (run under torchrun): gradients after GradReducer.finish() must equal the average
of the per-rank gradients of the same model on the ranks' own batches (developer tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn
from multimodal_ad_b200.models.Resnet3D import generate_model
from multimodal_ad_b200.sharding import GradReducer

def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = generate_model(model_depth=10, input_W=32, input_H=32, input_D=32, nb_class=3, pretrain_path=None, dropout_rate=0.0, device=dev)
    model.train()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.rand((2, 1, 32, 32, 32), device=dev, generator=g)
    y = torch.randint(0, 3, (2,), device=dev, generator=g)
    crit = nn.CrossEntropyLoss()
    # (a) no reducer: local gradients, averaged by hand
    crit(model(x), y).backward()
    want = []
    for p in model.parameters():
        t = p.grad.detach().clone() if p.grad is not None else None
        if t is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM); t /= world
        want.append(t)
    # running statistics changed in (a); gradients do not depend on them in training mode
    model.zero_grad(set_to_none=True)
    red = GradReducer(bucket_numel=1 << 20)
    model.grad_reducer = red
    crit(model(x), y).backward()
    red.finish(model.parameters())
    torch.cuda.synchronize()
    worst, nb = 0.0, 0
    for p, w in zip(model.parameters(), want):
        if w is None: continue
        worst = max(worst, ((p.grad - w).norm() / (w.norm() + 1e-20)).item()); nb += 1
    # gradient accumulation: a second backward WITHOUT zero_grad must add the averaged gradients once more
    crit(model(x), y).backward()
    red.finish(model.parameters())
    torch.cuda.synchronize()
    worst2 = 0.0
    for p, w in zip(model.parameters(), want):
        if w is None: continue
        worst2 = max(worst2, ((p.grad - 2 * w).norm() / (2 * w.norm() + 1e-20)).item())
    ok = worst < 1e-5 and worst2 < 1e-5
    if rank == 0:
        print(json.dumps(dict(world=world, params=nb, worst_rel=worst, worst_rel_accumulated=worst2, ok=ok)), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

if __name__ == "__main__":
    main()
