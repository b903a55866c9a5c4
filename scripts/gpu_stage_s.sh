#!/bin/bash
# stage S: packed-pair elementwise math + step splits in the decoder, splitter without hi write-back in the GEMM — gates, timing, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_tc.py tests/test_gpu_engine.py -m gpu -q --tb=short -rf -p no:cacheprovider --timeout 90 > gpurun_out/s_tests.log 2>&1
rc=$?; tail -8 gpurun_out/s_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "GATE FAILED rc=$rc"; grep -E "^FAILED|^E  " gpurun_out/s_tests.log | head -20 | cut -c1-200; exit 1; fi
timeout -k 5 100 python scripts/sym_tune.py 200000 2>&1 | tail -6
timeout -k 5 100 python scripts/sym_check.py 1000000 2>&1 | tail -7
timeout -k 5 120 python scripts/lab/sym_trace.py 200000 256 > gpurun_out/s_trace.log 2>&1; grep -E "==|EW group|issuer|flush|issued" gpurun_out/s_trace.log | cut -c1-330
timeout -k 5 200 python benchmarks/micro.py --only gemm 2>&1 | grep tf32x3 | cut -c1-250
timeout -k 5 600 python bench.py --steps 3 --warmup 3 --cpu-cells 4096 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/s_bench.json
