#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_scdeepsort.py tests/test_gpu_engine.py tests/test_gpu_scgnn2_module.py tests/test_gpu_em.py tests/test_gpu_examples.py -m gpu -q --tb=short -rf -p no:cacheprovider --timeout 120 > gpurun_out/x_tests.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/x_tests.log | cut -c1-300; grep -E "^FAILED|^E  " gpurun_out/x_tests.log | head -20 | cut -c1-250
