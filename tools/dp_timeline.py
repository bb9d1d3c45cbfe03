"""Kernel timeline of one data-parallel ResNet3D-18 training step (torchrun, one rank per GPU): where the NCCL all-reduces run
relative to the backward kernels.  Rank 0 records one eager step with torch.profiler (CUPTI kernel activity records: name,
stream, start, duration) and writes a compact summary: every NCCL kernel with the product kernels it overlaps, and the share
of all-reduce time hidden under compute.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dp_timeline.py OUT.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn
from torch.profiler import ProfilerActivity, profile

from multimodal_ad_b200.models.Resnet3D import generate_model
from multimodal_ad_b200.sharding import GradReducer


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dp_timeline.txt"
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = generate_model(model_depth=18, input_W=128, input_H=128, input_D=128, nb_class=3, pretrain_path=None, dropout_rate=0.5,
                           device=dev).train()
    reducer = GradReducer()
    model.grad_reducer = reducer
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-4, fused=True)
    crit = nn.CrossEntropyLoss()
    x = torch.rand(16, 1, 128, 128, 128, device=dev)
    y = torch.randint(0, 3, (16,), device=dev)

    def step():
        loss = crit(model(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        reducer.finish(model.parameters())
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):                                 # the first profiled steps absorb CUPTI's start-up skew between the ranks
            step()
            torch.cuda.synchronize()
            dist.barrier()
    dist.barrier()
    if rank == 0:
        ks = []
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None and "emcpy" not in e.name and "emset" not in e.name:
                ks.append((e.time_range.start, e.time_range.end, e.name))
        ks.sort()
        starts = [i for i, k in enumerate(ks) if "stem_s2d_pack" in k[2]]
        ks = ks[starts[-1]:]                               # the LAST profiled step
        t0 = ks[0][0]
        nccl = [k for k in ks if "nccl" in k[2].lower()]
        ours = [k for k in ks if "nccl" not in k[2].lower()]
        step_us = ks[-1][1] - t0
        lines = [f"# data parallel x{world}, rank 0, the last of 4 profiled eager ResNet3D-18 steps (batch 16 x 1x128^3): {len(ks)} kernels, {step_us / 1e3:.3f} ms "
                 f"from first kernel start to last kernel end (torch.profiler / CUPTI; profiling overhead included)",
                 "# NCCL kernels: start (ms into the step), duration (ms), share of the duration during which a product kernel runs, overlapping kernels"]
        hidden = total = 0.0
        for s, e, name in nccl:
            ov = [(max(s, a), min(e, b), n) for a, b, n in ours if a < e and b > s]
            # union of the overlap intervals
            cover, cur_s, cur_e = 0.0, None, None
            for a, b, _ in sorted(ov):
                if cur_e is None or a > cur_e:
                    if cur_e is not None:
                        cover += cur_e - cur_s
                    cur_s, cur_e = a, b
                else:
                    cur_e = max(cur_e, b)
            if cur_e is not None:
                cover += cur_e - cur_s
            total += e - s
            hidden += cover
            names = sorted({n.split("(")[0].replace("void ", "").replace("mmad::", "")[:40] for _, _, n in ov})
            lines.append(f"{(s - t0) / 1e3:8.3f} {(e - s) / 1e3:7.3f} {cover / max(e - s, 1e-9):5.2f}  {name.split('(')[0][:48]:48s} | {', '.join(names[:6])}")
        lines.insert(1, f"# all-reduce kernel time {total / 1e3:.3f} ms in {len(nccl)} kernels, {100 * hidden / max(total, 1e-9):.1f} % of it concurrent with product kernels; "
                        f"last NCCL kernel ends {(max(e for _, e, _ in nccl) - t0) / 1e3:.3f} ms into the step" if nccl else "# no NCCL kernels recorded")
        with open(out_path, "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines[:12]))
    dist.destroy_process_group()


main()
