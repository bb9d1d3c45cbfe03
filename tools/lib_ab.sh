#!/bin/bash
# A/B of two builds of the library on the headline step (usage: tools/lib_ab.sh <other.so>)
for rep in 1 2; do
for lib in "" "$1"; do
  MMAD_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/ab.log 2> gpurun_out/ab.err
  python -c "import json; d=json.loads(open('gpurun_out/ab.log').read().strip().split(chr(10))[-1]); print('lib=${lib:-current}', round(d['ms_per_step'],3), round(d['value'],1))" || tail -5 gpurun_out/ab.err
done
done
