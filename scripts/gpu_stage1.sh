#!/bin/bash
# Stage-1 GPU check of round 2: the new symmetric decoder first (own short timeout: a deadlocked kernel must not hang the box),
# then the full GPU suite with the decoder forced to the row-sweep path if the symmetric tests failed, then the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "symmetric" > gpurun_out/s1_sym.log 2>&1
echo "sym rc=$?" >> gpurun_out/s1_sym.log
tail -25 gpurun_out/s1_sym.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/s1_all.log 2>&1
echo "all rc=$?" >> gpurun_out/s1_all.log
tail -40 gpurun_out/s1_all.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench rc=$?"
head -c 4000 gpurun_out/s1_bench.json; tail -c 2000 gpurun_out/s1_bench.err
