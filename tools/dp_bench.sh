#!/bin/bash
# data-parallel headline step at N GPUs: graph vs eager, NCCL CTA caps  (usage: tools/dp_bench.sh N)
N=${1:-2}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/dp_${N}_$name.log 2> gpurun_out/dp_${N}_$name.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dp_${N}_$name.log").read().strip().split("\n")[-1])
    print(d["value"], d["ms_per_step"], d["config"]["launch"], d["e2e"]["value"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/dp_${N}_$name.err").read()[-1500:])
PY
}
run graph X=1
run eager MMAD_BENCH_DP_GRAPH=0
run graph_cta8 NCCL_MAX_CTAS=8
run graph_cta4 NCCL_MAX_CTAS=4
