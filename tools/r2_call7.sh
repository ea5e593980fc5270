#!/bin/bash
# host-side profile of the timed loop (cProfile on rank 0), 1 GPU and N GPUs
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 200 python bench.py --no-e2e --no-cpu --steps 40 --profile > gpurun_out/c7_prof1.log 2>&1; echo "prof1 rc=$?"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --no-e2e --no-cpu --steps 40 --profile > gpurun_out/c7_prof$N.log 2>&1; echo "prof$N rc=$?"
OA_EXCHANGE_BATCH=16 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --no-e2e --no-cpu --steps 48 --profile > gpurun_out/c7_prof${N}_k16.log 2>&1; echo "prof${N}k16 rc=$?"
