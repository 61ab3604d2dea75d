#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/bringup_pair.sh > /dev/null 2>&1
grep -c '"ok": true' gpurun_out/bringup_pair.txt; grep -B2 "rc=[1-9]" gpurun_out/bringup_pair.txt | head -20
if grep -q "rc=[1-9]" gpurun_out/bringup_pair.txt; then echo "BRINGUP FAILURES"; exit 0; fi
for m in 0 2; do echo "== LOFT_2CTA=$m"; LOFT_2CTA=$m timeout -s KILL 120 python tools/gemm_timeline.py 2>&1 | tail -13; done | tee gpurun_out/gemm_timeline_e.txt
for m in 0 2 1; do
  LOFT_2CTA=$m timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes_v2_2cta$m.txt 2>&1
  echo "gemm_shapes mode $m rc=$?"; tail -1 gpurun_out/gemm_shapes_v2_2cta$m.txt
done
