#!/bin/bash
# stage N: warp-uniform MMA issue in the symmetric decoder — correctness gate, timing, trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 90 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "symmetric or large_embedding or pair_sharded" --tb=short -rf -p no:cacheprovider --timeout 60 > gpurun_out/n_tests.log 2>&1
rc=$?; tail -5 gpurun_out/n_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "GATE FAILED rc=$rc"; exit 1; fi
timeout -k 5 100 python scripts/sym_tune.py 200000 2>&1 | tail -6
timeout -k 5 100 python scripts/sym_check.py 1000000 2>&1 | tail -7
timeout -k 5 120 python scripts/lab/sym_trace.py 200000 256 > gpurun_out/n_trace.log 2>&1; head -8 gpurun_out/n_trace.log | cut -c1-330
