#!/bin/bash
# One `gpurun --gpus 2` call: NCCL gradient parity of the sharded loss head on 2 GPUs (tools/dist_check.py, the test
# the 1-GPU suite skips) and the 2-GPU bench line (dist_parity checked before timing).
set -x
cd "$(dirname "$0")/.."
R=${ROUND:-r02d}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/dist_check.py > gpurun_out/${R}_dist_check_n2.log 2>&1; tail -n 4 gpurun_out/${R}_dist_check_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err; tail -c 1500 gpurun_out/${R}_bench_n2.json
