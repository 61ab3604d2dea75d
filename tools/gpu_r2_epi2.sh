#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -2
echo "== timeline (product build)"; LOFT_2CTA=0 timeout -s KILL 120 python tools/gemm_timeline.py p2 l3 l4 l3_1x1 l4_1x1 fc2 2>&1 | tail -6
echo "== ktrace build, epilogue cycles"; LOFT_LIB_PATH=$PWD/build/ktrace/libloft_b200_ktrace.so LOFT_2CTA=0 timeout -s KILL 120 python tools/gemm_timeline.py p2 l3 l3_1x1 fc2 2>&1 | grep -o "^[a-z0-9_]* \|epilogue of the first tile.*" | paste - -
timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes_epi2.txt 2>&1; tail -1 gpurun_out/gemm_shapes_epi2.txt
timeout -s KILL 100 python tools/probe_epi.py res; timeout -s KILL 100 python tools/epi_bench.py 2>&1 | tail -12
} 2>&1 | tee gpurun_out/r02_epi2.txt
