#!/bin/bash
# usage: tools/run_mgpu.sh N TAG [bench args] -- multi-rank parity worker + bench at N GPUs (under gpurun --gpus N)
N=${1:-2}; TAG=${2:-x}; shift; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py > gpurun_out/mgpu_${TAG}.log 2>&1
echo "mgpu_worker rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/mgpu_${TAG}.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err
echo "bench rc=$?"; python - <<PY
import json
for l in open("gpurun_out/bench_n${N}_${TAG}.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["kernels"], d["gpu_launches"], d.get("lanczos_gs"), d["e2e"])
PY
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n${N}_${TAG}.err | tail -5
