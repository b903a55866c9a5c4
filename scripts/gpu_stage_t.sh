#!/bin/bash
# stage T: round-2 profile evidence — launch list of one timed step, ncu --set full of the decoder, the aggregate and the GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step_1m.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-checks --cuda-profiler > gpurun_out/t_launch_bench.json 2> gpurun_out/t_launch_bench.err
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_launches_step_1m.csv)"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:gae_sym_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_gae_sym_1m -f \
    python scripts/sym_one.py 1000000 sym > gpurun_out/t_ncu_sym.log 2>&1
echo "ncu sym rc=$?"
for v in f32 bf16; do
  timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:spmm_stream --launch-skip 1 --launch-count 1 \
      -o gpurun_out/r02_spmm_stream_1m_$v -f python scripts/spmm_probe.py 1000000 32 1 $v locality > gpurun_out/t_ncu_spmm_$v.log 2>&1
  echo "ncu spmm $v rc=$?"
done
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 40 --launch-count 2 -o gpurun_out/r02_gemm_tc -f \
    python bench.py --cells 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-checks > gpurun_out/t_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -6
