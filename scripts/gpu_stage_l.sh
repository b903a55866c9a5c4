#!/bin/bash
# stage L: aggregate lab, generation 2 (pair stream staged through shared memory) + ncu of two variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python scripts/spmm_probe.py 1000000 32 3 f32 locality /tmp/csr.bin > gpurun_out/l_probe.log 2>&1
echo "probe rc=$?"; tail -2 gpurun_out/l_probe.log
: > gpurun_out/l_lab.log
for v in $(./scripts/lab/gather_lab x list | grep -E "^s2:|^st:f32:l:16:2|^st:bf16:l:4:3"); do
  timeout -k 5 60 ./scripts/lab/gather_lab /tmp/csr.bin "$v" 5 >> gpurun_out/l_lab.log 2>&1
  echo "rc=$? $v" >> gpurun_out/l_lab.log
done
grep -E "cold|TIMEOUT|failed|rror|rc=[1-9]" gpurun_out/l_lab.log
for v in s2:f32:16:2:8 s2:bf16:4:3:8; do
  timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:"stream2" --launch-skip 1 --launch-count 1 \
      -o gpurun_out/r02_lab_$(echo $v | tr ':' '_') -f ./scripts/lab/gather_lab /tmp/csr.bin $v 2 > gpurun_out/l_ncu.log 2>&1
  echo "ncu $v rc=$?"
done
