#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/profile_run.log 2>&1
echo "ncu rc=$?" >> gpurun_out/profile_run.log
tail -3 gpurun_out/profile_run.log
python tools/summarize_ncu.py gpurun_out/launches.csv gpurun_out/launch_summary.txt | head -60
