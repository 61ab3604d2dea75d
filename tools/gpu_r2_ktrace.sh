#!/bin/bash
# cycle accounting of the k-loop (instrumented build: make with -DLOFT_KTRACE into build/ktrace)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LOFT_LIB_PATH=$PWD/build/ktrace/libloft_b200_ktrace.so
{
for k in 1 2; do for c in 0 2; do echo "== LOFT_KGROUP=$k LOFT_2CTA=$c"; LOFT_KGROUP=$k LOFT_2CTA=$c timeout -s KILL 120 python tools/gemm_timeline.py p2 l3 l4 fc1 l3_2d l4_2d foa_g4 2>&1 | tail -14; done; done
for s in 1 2 3 7; do echo "== LOFT_GEMM_SKIP=$s KGROUP=1 1-CTA"; LOFT_KGROUP=1 LOFT_2CTA=0 LOFT_GEMM_SKIP=$s timeout -s KILL 120 python tools/gemm_timeline.py l3_2d l4_2d 2>&1 | tail -4; done
} 2>&1 | tee gpurun_out/r02_ktrace.txt
