#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 300 python scripts/diag_bench_nan.py > gpurun_out/e_diag.log 2>&1; echo "diag rc=$?"; tail -n 120 gpurun_out/e_diag.log | cut -c1-200
timeout -k 5 120 python scripts/sym_check.py 1000000 > gpurun_out/e_sym.log 2>&1; tail -n 8 gpurun_out/e_sym.log
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:gae_sym_kernel -c 1 -o gpurun_out/r02_sym_200k python scripts/sym_one.py 200000 > gpurun_out/e_ncu.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/e_ncu.log
