#!/bin/bash
# usage: tools/sweep_env.sh N VAR v1 v2 ... -- bench (no Lanczos / CPU baseline) at N GPUs for each value of env VAR
N=$1; VAR=$2; shift; shift
for c in "$@"; do
  echo "== $VAR=$c"
  env $VAR=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --lanczos 0 --e2e-steps 1 --no-cpu-baseline $BENCH_ARGS 2>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], [round(k['ms'],4) for k in d['kernels']])
"
done
