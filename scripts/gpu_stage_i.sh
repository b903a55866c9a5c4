#!/bin/bash
# stage I: 2-GPU bench (pair-sharded symmetric decoder + dz all-reduce), torch.distributed and C-ABI NCCL communicators
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for comm in torch native; do
  timeout -k 5 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM % 10)) \
      bench.py --gpus 2 --steps 3 --warmup 3 --comm $comm --no-cpu-baseline > gpurun_out/i_bench2_$comm.json 2> gpurun_out/i_bench2_$comm.err
  echo "comm=$comm rc=$?"
  tail -c 600 gpurun_out/i_bench2_$comm.json
  tail -5 gpurun_out/i_bench2_$comm.err
done
