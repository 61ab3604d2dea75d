#!/bin/bash
# round-2 GPU visit A: pair-mode bring-up, then per-shape A/B timing (1-CTA vs forced pair vs policy)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/bringup_pair.sh > /dev/null 2>&1
grep -B1 -A1 "rc=" gpurun_out/bringup_pair.txt | grep -v "^--" | paste - - - | awk '{print}' | head -80
if grep -A2 "LOFT_2CTA=2" gpurun_out/bringup_pair.txt | grep -q "rc=[1-9]"; then
  echo "PAIR MODE FAILURES -- skipping timing"; exit 0
fi
for m in 0 2 1; do
  LOFT_2CTA=$m timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes_2cta$m.txt 2>&1
  echo "gemm_shapes mode $m rc=$?"; tail -1 gpurun_out/gemm_shapes_2cta$m.txt
done
