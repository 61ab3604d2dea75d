#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:loft_gemm_tf32_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_epi python tools/probe_epi.py res > gpurun_out/r02_prof_epi.log 2>&1; echo rc=$?
tail -3 gpurun_out/r02_prof_epi.log
