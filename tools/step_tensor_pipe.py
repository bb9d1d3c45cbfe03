"""Step-level tensor-pipe utilisation from an ncu launch list that carries, per launch, gpu__time_duration.sum and
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed (csv, one row per launch and metric):

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
        --clock-control none -c 1200 --csv --log-file gpurun_out/step_pipe.csv python tools/resnet_bench.py ...
    python tools/step_tensor_pipe.py gpurun_out/step_pipe.csv profiles/r02_resnet_step_tensor_pipe.json [profiles/r02_resnet_launches.txt]

The last training step of the capture (from its stem_s2d_pack launch on) is summarised: time-weighted mean of the per-kernel
tensor-pipe activity over ALL launches of the step (ncu times are cold-cache and serialised, so this is a share-weighted figure,
not a wall-clock one), plus the same over the tensor-core kernels only."""
import collections
import csv
import json
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        if not r[idi].isdigit():
            continue
        d = launches.setdefault(int(r[idi]), {"name": r[ki]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    seq = list(launches.values())
    starts = [i for i, d in enumerate(seq) if "stem_s2d_pack" in d["name"]]
    # the last step of the capture when it is complete (the first step also carries the optimizer's one-off state initialisation:
    # ~190 fill kernels), else the one before it
    step = seq[starts[-1]:]
    complete = any("multi_tensor_apply" in d["name"] for d in step[-12:])      # ends with the fused Adam update
    if len(starts) >= 2 and not complete:
        step = seq[starts[-2]:starts[-1]]
    tkey = "gpu__time_duration.sum"
    pkey = [k for k in step[0] if k.startswith("sm__pipe_tensor_cycles_active")][0]
    tot = sum(d[tkey] for d in step)
    weighted = sum(d[tkey] * d.get(pkey, 0.0) for d in step) / tot
    tc = [d for d in step if re.search(r"conv3d|stem_conv|stem_wgrad", d["name"])]
    tc_t = sum(d[tkey] for d in tc)
    out = {"metric": pkey, "launches": len(step), "step_time_ncu_ms": tot / 1e6, "step_weighted_pct": weighted,
           "tensor_kernels_share_of_time": tc_t / tot, "tensor_kernels_weighted_pct": sum(d[tkey] * d.get(pkey, 0.0) for d in tc) / tc_t,
           "how": "ncu --clock-control none, per-launch values weighted by per-launch gpu__time_duration over one eager training step "
                  "(batch 16 x 1x128^3, tools/resnet_bench.py)"}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))
    if len(sys.argv) > 3:                                  # the per-kernel table of the same step (profiles/rNN_resnet_launches.txt)
        agg = collections.OrderedDict()
        for d in step:
            name = re.sub(r"\(.*", "", re.sub(r"^void ", "", d["name"])).replace("mmad::", "")[:64]
            a = agg.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += d[tkey]
            a[2] += d[tkey] * d.get(pkey, 0.0)
        with open(sys.argv[3], "w") as f:
            f.write("# one eager ResNet3D-18 training step (batch 16 x 1x128^3), ncu --clock-control none; per kernel: summed "
                    "gpu__time_duration, share, launches, time-weighted sm__pipe_tensor_cycles_active (% of peak, elapsed)\n")
            f.write(f"# {len(step)} launches, {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares)\n")
            for name, (cnt, t, tp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{t / 1e6:8.3f} ms {100 * t / tot:5.1f}% x{cnt:3d}  tensor pipe {tp / t if t else 0.0:5.1f}%  {name}\n")


main()
