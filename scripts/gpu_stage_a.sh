#!/bin/bash
# short, verbose, hard-killed stages: a deadlocked kernel must never eat the box's time limit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; secs=$2; shift 2; timeout -k 5 "$secs" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name rc=$?" | tee -a "gpurun_out/$name.log"; tail -n "${TAILN:-30}" "gpurun_out/$name.log" | cut -c1-400; }
run a_sym 150 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "symmetric" --tb=short -rf -p no:cacheprovider
run a_em 400 python -m pytest tests/test_gpu_em.py tests/test_gpu_prep.py tests/test_gpu_layers.py -m gpu -q --tb=short -rf --timeout 120 -p no:cacheprovider
