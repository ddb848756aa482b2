#!/bin/bash
# usage: tools/sweep_env.sh N "VAR=a VAR2=b" "VAR=c" ... -- bench (no Lanczos / CPU baseline / parity) at N GPUs per env set
N=$1; shift
for c in "$@"; do
  echo "== $c"
  env $c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --lanczos 0 --e2e-steps 1 --no-cpu-baseline --no-parity --no-cfg3 --cfg4 0 $BENCH_ARGS 2>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], [round(k['ms'],4) for k in d['kernels']])
"
done
