#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 900 python -m pytest tests -m gpu -q --tb=short -rf --timeout 240 -p no:cacheprovider > gpurun_out/h_all.log 2>&1; echo "all rc=$?"; tail -n 30 gpurun_out/h_all.log | cut -c1-400
timeout -k 5 400 python benchmarks/configs.py --only 1,3,5 > gpurun_out/h_configs.jsonl 2> gpurun_out/h_configs.err; echo "configs rc=$?"; cat gpurun_out/h_configs.jsonl | cut -c1-700; tail -c 600 gpurun_out/h_configs.err
