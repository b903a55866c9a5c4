#!/bin/bash
# gated stages with hard kills: symmetric decoder → new rows → full GPU suite → bench (each only if the previous gate passed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; secs=$2; shift 2; timeout -k 5 "$secs" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "$name rc=$rc" | tee -a "gpurun_out/$name.log"; tail -n "${TAILN:-25}" "gpurun_out/$name.log" | cut -c1-500; return $rc; }
run b_sym 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "symmetric or single_column or tensor_core_path" --tb=short -rf -p no:cacheprovider --timeout 90 || { echo "GATE: symmetric decoder failed; forcing the row-sweep path for the rest"; export B2_FORCE_GAE_PATH=f16; }
run b_new 400 python -m pytest tests/test_gpu_em.py tests/test_gpu_prep.py tests/test_gpu_layers.py tests/test_gpu_examples.py -m gpu -q --tb=short -rf --timeout 150 -p no:cacheprovider
TAILN=40 run b_all 900 python -m pytest tests -m gpu -q --tb=line -rf --timeout 200 -p no:cacheprovider --deselect tests/test_gpu_em.py --deselect tests/test_gpu_prep.py --deselect tests/test_gpu_layers.py --deselect tests/test_gpu_examples.py
timeout -k 5 500 python bench.py --steps 5 --warmup 3 --cpu-cells 4096 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
head -c 6000 gpurun_out/b_bench.json; tail -c 1500 gpurun_out/b_bench.err
