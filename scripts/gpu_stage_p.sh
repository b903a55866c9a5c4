#!/bin/bash
# stage P: two-warp MMA issue (decoder) + elected issue in the GEMM — correctness gates, timing, trace; then the aggregate lab gen 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 150 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_tc.py -m gpu -q -k "symmetric or large_embedding or pair_sharded or gemm" --tb=short -rf -p no:cacheprovider --timeout 60 > gpurun_out/p_tests.log 2>&1
rc=$?; tail -5 gpurun_out/p_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "GATE FAILED rc=$rc"; exit 1; fi
timeout -k 5 100 python scripts/sym_tune.py 200000 2>&1 | tail -6
timeout -k 5 100 python scripts/sym_check.py 1000000 2>&1 | tail -7
timeout -k 5 120 python scripts/lab/sym_trace.py 200000 256 > gpurun_out/p_trace.log 2>&1; grep -E "==|EW group|issuer|flush|issued" gpurun_out/p_trace.log | cut -c1-330
timeout -k 5 200 python benchmarks/micro.py --only gemm 2>&1 | tail -12 | cut -c1-250
bash scripts/gpu_stage_o.sh
