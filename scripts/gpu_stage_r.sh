#!/bin/bash
# stage R: conversion-free fp16 planes in the decoder's elementwise math; lean aggregate epilogue
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "symmetric or large_embedding or pair_sharded or spmm or gae or step_splits" --tb=short -rf -p no:cacheprovider --timeout 90 > gpurun_out/r_tests.log 2>&1
rc=$?; tail -8 gpurun_out/r_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "GATE FAILED rc=$rc"; exit 1; fi
timeout -k 5 100 python scripts/sym_tune.py 200000 2>&1 | tail -6
timeout -k 5 100 python scripts/sym_check.py 1000000 2>&1 | tail -7
timeout -k 5 120 python scripts/lab/sym_trace.py 200000 256 > gpurun_out/r_trace.log 2>&1; grep -E "==|EW group|issuer|flush|issued" gpurun_out/r_trace.log | cut -c1-330
timeout -k 5 200 python scripts/spmm_probe.py 1000000 32 5 f32,bf16 locality 2>&1 | tail -2
