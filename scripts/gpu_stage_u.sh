#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 200 python scripts/sym_splits.py 1000000 1,2,0,3 2>&1 | tail -5
timeout -k 5 100 python scripts/sym_splits.py 500000 1,2,4,0 2>&1 | tail -5
timeout -k 5 100 python scripts/sym_splits.py 125000 1,2,4,8,0 2>&1 | tail -6
timeout -k 5 500 python benchmarks/configs.py --only 1,3,5 > gpurun_out/u_configs.jsonl 2> gpurun_out/u_configs.err; cat gpurun_out/u_configs.jsonl | cut -c1-400
