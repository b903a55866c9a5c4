#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 100 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_tc.py -m gpu -q -k "symmetric or large_embedding or tensor_core or gemm" --tb=short -rf -p no:cacheprovider --timeout 90 2>&1 | tail -6 | cut -c1-300
timeout -k 5 100 python scripts/sym_tune.py 200000 2>&1 | tail -12
timeout -k 5 100 python scripts/sym_check.py 1000000 2>&1 | tail -7
