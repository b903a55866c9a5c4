#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 120 python scripts/lab/sym_trace.py 200000 256 2>&1 | tail -60
