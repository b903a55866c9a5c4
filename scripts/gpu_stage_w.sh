#!/bin/bash
# stage W: 4-GPU sanity run of the bench (pair-sharded decoder with step splits, row-sharded aggregate)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench4.json 2> gpurun_out/w_bench4.err
echo "rc=$?"; tail -c 700 gpurun_out/w_bench4.json; grep -n "Error\|error" gpurun_out/w_bench4.err | head -5
