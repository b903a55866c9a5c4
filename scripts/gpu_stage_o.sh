#!/bin/bash
# stage O: aggregate lab, generation 3 (row slabs for L2 residency + L2 eviction hints)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python scripts/spmm_probe.py 1000000 32 3 f32 locality /tmp/csr.bin > gpurun_out/o_probe.log 2>&1
echo "probe rc=$?"; tail -1 gpurun_out/o_probe.log
: > gpurun_out/o_lab.log
for v in $(./scripts/lab/gather_lab x list | grep -E "^s3:|^s2:f32:16:2:8|^s2:bf16:4:2:8"); do
  timeout -k 5 60 ./scripts/lab/gather_lab /tmp/csr.bin "$v" 5 >> gpurun_out/o_lab.log 2>&1
  echo "rc=$? $v" >> gpurun_out/o_lab.log
done
grep -E "cold|TIMEOUT|failed|rror|rc=[1-9]" gpurun_out/o_lab.log
for v in s3:f32:16:2:8:h1:s10; do
  timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:"stream3" --launch-skip 12 --launch-count 1 \
      -o gpurun_out/r02_lab_$(echo $v | tr ':' '_') -f ./scripts/lab/gather_lab /tmp/csr.bin $v 2 > gpurun_out/o_ncu.log 2>&1
  echo "ncu $v rc=$?"
done
