#!/bin/bash
# stage V: final validation — full GPU suite, smoke(), default bench + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 1000 python -m pytest tests -m gpu -q --tb=short -rf --timeout 240 -p no:cacheprovider > gpurun_out/v_all.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/v_all.log | cut -c1-300
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 5 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/v_bench_ref.json
timeout -k 5 600 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/v_bench.json
