#!/bin/bash
# stage J: aggregate lab — real kNN graph at 1 M cells, every gather mechanism (each variant in its own process: a faulting
# experimental kernel must not take the others down); then ncu --set full of the two library kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python scripts/spmm_probe.py 1000000 32 5 f32,bf16 locality /tmp/csr.bin > gpurun_out/j_probe.log 2>&1
echo "probe rc=$?"; tail -3 gpurun_out/j_probe.log
: > gpurun_out/j_lab.log
for v in $(./scripts/lab/gather_lab x list); do
  timeout -k 5 60 ./scripts/lab/gather_lab /tmp/csr.bin "$v" 5 >> gpurun_out/j_lab.log 2>&1
  echo "rc=$? $v" >> gpurun_out/j_lab.log
done
grep -E "cold|TIMEOUT|failed|error|rc=[1-9]" gpurun_out/j_lab.log
for v in f32 bf16; do
  timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:"spmm_csr" --launch-skip 1 --launch-count 1 \
      -o gpurun_out/r02_spmm_1m_$v -f python scripts/spmm_probe.py 1000000 32 1 $v locality > gpurun_out/j_ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done
