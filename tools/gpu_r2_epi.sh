#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LOFT_LIB_PATH=$PWD/build/ktrace/libloft_b200_ktrace.so
{
for s in 0 8 16 24; do echo "== LOFT_GEMM_SKIP=$s (8 no stores, 16 no tmem loads)"; LOFT_2CTA=0 LOFT_GEMM_SKIP=$s timeout -s KILL 120 python tools/gemm_timeline.py p2 l3 l4 l3_1x1 fc2 2>&1 | tail -10; done
} 2>&1 | tee gpurun_out/r02_epi_attr.txt
