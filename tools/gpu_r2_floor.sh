#!/bin/bash
# Attribution of the k-loop floor (~200 ns per k-block at <= 128 columns): which resource is it?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in 0 1 2 3 4 7; do echo "== LOFT_GEMM_SKIP=$s (1 no A, 2 no B, 4 one MMA)"; LOFT_2CTA=0 LOFT_GEMM_SKIP=$s timeout -s KILL 120 python tools/gemm_timeline.py l3_2d l4_2d fc1 2>&1 | tail -3; done
for m in 8 32 64 148; do echo "== LOFT_GEMM_MAXCTAS=$m"; LOFT_2CTA=0 LOFT_GEMM_MAXCTAS=$m timeout -s KILL 120 python tools/gemm_timeline.py l3 l4 l3_2d l4_2d fc1 2>&1 | tail -5; done
for n in 64 128 256; do echo "== LOFT_TILE_N=$n"; LOFT_2CTA=0 LOFT_TILE_N=$n timeout -s KILL 120 python tools/gemm_timeline.py l3 l4 l3_2d l4_2d 2>&1 | tail -4; done
for n in 64 128 256; do echo "== pair LOFT_TILE_N=$n"; LOFT_2CTA=2 LOFT_TILE_N=$n timeout -s KILL 120 python tools/gemm_timeline.py l3 l4 l3_2d l4_2d 2>&1 | tail -4; done
} 2>&1 | tee gpurun_out/r02_floor.txt
