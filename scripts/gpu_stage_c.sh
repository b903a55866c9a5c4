#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; secs=$2; shift 2; timeout -k 5 "$secs" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "$name rc=$rc" | tee -a "gpurun_out/$name.log"; tail -n "${TAILN:-25}" "gpurun_out/$name.log" | cut -c1-500; return $rc; }
run c_sym 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "symmetric or single_column or tensor_core_path" --tb=short -rf -p no:cacheprovider --timeout 90 || exit 1
TAILN=16 run c_diag 240 python scripts/sym_check.py 200000 1000000
run c_few 400 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_scgnn2_module.py tests/test_gpu_em.py tests/test_gpu_examples.py -m gpu -q --tb=short -rf --timeout 200 -p no:cacheprovider
