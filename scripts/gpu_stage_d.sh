#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 120 python -m pytest tests/test_gpu_scgnn2_module.py tests/test_gpu_full_size.py -m gpu -q --tb=short -rf --timeout 100 -p no:cacheprovider 2>&1 | tail -15 | cut -c1-300
timeout -k 5 600 python bench.py --steps 5 --warmup 3 --cpu-cells 4096 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
head -c 7000 gpurun_out/d_bench.json; tail -c 1800 gpurun_out/d_bench.err
