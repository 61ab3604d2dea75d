#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -2
for c in 0 2; do echo "== LOFT_2CTA=$c"; LOFT_2CTA=$c timeout -s KILL 120 python tools/gemm_timeline.py 2>&1 | tail -14; done
echo "== ktrace build"; LOFT_LIB_PATH=$PWD/build/ktrace/libloft_b200_ktrace.so LOFT_2CTA=0 timeout -s KILL 120 python tools/gemm_timeline.py p2 l3 l4 fc1 l3_2d 2>&1 | tail -10
timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes_kloop2.txt 2>&1; tail -1 gpurun_out/gemm_shapes_kloop2.txt
} 2>&1 | tee gpurun_out/r02_kloop2.txt
