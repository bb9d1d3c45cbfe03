#!/bin/bash
# torchrun bench at N GPUs (usage: tools/scale_check.sh N [extra bench flags])
N=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$N.log").read().strip().split("\n")[-1])
    print("N=$N", round(d["value"],1), round(d["ms_per_step"],3), d["config"]["launch"], "e2e", round(d["e2e"]["value"],1))
    for k in ("roi_pool","resnet3d18_train_91x109x91","resnet3d50_train","unet3d_roi_extract"):
        if k in d: print("  ", k, d[k].get("value"), d[k].get("ms_per_step"), d[k].get("error"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/scale_$N.err").read()[-2000:])
PY
