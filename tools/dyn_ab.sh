#!/bin/bash
# A/B of the dynamic tile scheduler at N GPUs (usage: tools/dyn_ab.sh N)
N=${1:-2}
for d in 1 0; do
  if [ "$N" = "1" ]; then
    MMAD_CONV_DYN=$d timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/ab.log 2> gpurun_out/ab.err
  else
    MMAD_CONV_DYN=$d timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/ab.log 2> gpurun_out/ab.err
  fi
  python -c "import json; d=json.loads(open('gpurun_out/ab.log').read().strip().split(chr(10))[-1]); print('N=$N dyn=$d', round(d['ms_per_step'],3), round(d['value'],1), d['config']['launch'])" || tail -5 gpurun_out/ab.err
done
