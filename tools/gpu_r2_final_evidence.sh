#!/bin/bash
# round-2 final evidence: launch list of one bench step + ncu --set full of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r02f_profile_run.log 2>&1
echo "launch list rc=$?"
python tools/summarize_ncu.py gpurun_out/r02f_launches.csv gpurun_out/r02f_launch_summary.txt | head -12
full() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 ${@:5} > gpurun_out/$4.log 2>&1; echo "$4 rc=$?"; }
full loft_gemm_tf32_kernel 9 3 r02f_prof_foa python tools/probe_foa.py 105
full loft_gemm_tf32_kernel 3 1 r02f_prof_l4 python tools/probe_l4.py
full loft_gemm_tf32_kernel 3 1 r02f_prof_p2 python tools/probe_p2.py
full loft_gemm_tf32_kernel 4 1 r02f_prof_epi python tools/probe_epi.py res
full roi_align_kernel 6 2 r02f_prof_roi python tools/roi_bench.py
full soft_nms_fast_kernel 2 1 r02f_prof_softnms python tools/soft_nms_bench.py
full poly_ 2 2 r02f_prof_poly python tools/pipeline_bench.py
python tools/ncu_rep_summary.py gpurun_out/r02f_prof_*.ncu-rep > gpurun_out/r02f_ncu_kernels.txt 2>&1; wc -l gpurun_out/r02f_ncu_kernels.txt
