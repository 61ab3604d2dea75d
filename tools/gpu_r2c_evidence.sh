#!/bin/bash
# round-2 (second half) evidence: launch list of a bench step + ncu --set full of the new kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02c_launches.csv \
    python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r02c_profile_run.log 2>&1
echo "launch list rc=$?"
python tools/summarize_ncu.py gpurun_out/r02c_launches.csv gpurun_out/r02c_launch_summary.txt | head -14
full() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 ${@:5} > gpurun_out/$4.log 2>&1; echo "$4 rc=$?"; }
full "stem_pack|loft_gemm_tf32" 4 2 r02c_prof_stem python tools/probe_r2c.py stem
full "rpn_hist|rpn_gather|rpn_write|iou_" 10 5 r02c_prof_rpn python tools/probe_r2c.py rpn
full "narrow_" 4 2 r02c_prof_narrow python tools/probe_r2c.py narrow
full "gather_rot|scatter_rot" 4 2 r02c_prof_rot python tools/probe_r2c.py rot
python tools/ncu_rep_summary.py gpurun_out/r02c_prof_*.ncu-rep > gpurun_out/r02c_ncu_kernels.txt 2>&1; wc -l gpurun_out/r02c_ncu_kernels.txt
