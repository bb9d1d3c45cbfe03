#!/bin/bash
N=${1:-2}
for b in 8388608 4194304 2097152 8388608; do
  MMAD_BUCKET_NUMEL=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/ab.log 2> gpurun_out/ab.err
  python -c "import json; d=json.loads(open('gpurun_out/ab.log').read().strip().split(chr(10))[-1]); print('N=$N bucket=$b', round(d['ms_per_step'],3), round(d['value'],1))" || tail -5 gpurun_out/ab.err
done
