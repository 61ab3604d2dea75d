#!/bin/bash
# round-2 evidence: launch list of one bench step + ncu --set full of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r02_profile_run.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r02_profile_run.log | cut -c1-300
python tools/summarize_ncu.py gpurun_out/r02_launches.csv gpurun_out/r02_launch_summary.txt | head -40
full() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 ${@:5} > gpurun_out/$4.log 2>&1; echo "$4 rc=$?"; }
full loft_gemm_tf32_kernel 9 3 r02_prof_foa python tools/probe_foa.py 105
full loft_gemm_tf32_kernel 3 1 r02_prof_l4 python tools/probe_l4.py
full loft_gemm_tf32_kernel 3 1 r02_prof_p2 python tools/probe_p2.py
full roi_align_kernel 6 2 r02_prof_roi python tools/roi_bench.py
ls -la gpurun_out/*.ncu-rep
