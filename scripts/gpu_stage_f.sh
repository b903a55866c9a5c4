#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "symmetric or large_embedding or single_column or tensor_core_path" --tb=short -rf -p no:cacheprovider --timeout 90 2>&1 | tail -12 | cut -c1-300
timeout -k 5 150 python scripts/sym_tune.py 200000 2>&1 | tail -14
timeout -k 5 600 python bench.py --steps 5 --warmup 3 --cpu-cells 4096 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
head -c 9000 gpurun_out/f_bench.json; tail -c 1500 gpurun_out/f_bench.err
