#!/bin/bash
# k-blocks per barrier slot: parity of the kernel tests, then the per-CTA timeline and every shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
for k in 1 2; do for c in 0 2; do echo "== LOFT_KGROUP=$k LOFT_2CTA=$c"; LOFT_KGROUP=$k LOFT_2CTA=$c timeout -s KILL 120 python tools/gemm_timeline.py 2>&1 | tail -14; done; done
for k in 1 2; do
  LOFT_KGROUP=$k timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes_kgroup$k.txt 2>&1
  echo "gemm_shapes kgroup $k rc=$?"; tail -1 gpurun_out/gemm_shapes_kgroup$k.txt
done
} 2>&1 | tee gpurun_out/r02_kgroup.txt
