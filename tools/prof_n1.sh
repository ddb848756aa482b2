#!/bin/bash
# usage (under gpurun, 1 GPU): tools/prof_n1.sh TAG -- bench without ncu, then the launch list, then one --set full capture
TAG=${1:-r2}
ARGS="--steps 3 --warmup 3 --lanczos 0 --no-cpu-baseline --e2e-steps 1"
python bench.py $ARGS > gpurun_out/prof_${TAG}_plain.json 2> gpurun_out/prof_${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $ARGS > gpurun_out/prof_${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fastc|k_slow" -s 4 -c 2 -f -o gpurun_out/ncu_${TAG} python bench.py $ARGS > gpurun_out/prof_${TAG}_ncu2.log 2>&1
ls -la gpurun_out/ncu_${TAG}.ncu-rep; tail -2 gpurun_out/prof_${TAG}_ncu2.log
