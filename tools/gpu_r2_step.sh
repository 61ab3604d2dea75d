#!/bin/bash
# whole-step check after a kernel change: GPU tests, bench line, idle-gap timeline, phase times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-step}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python -c "import sys,json; d=[json.loads(l) for l in open('gpurun_out/bench_$TAG.json') if l.startswith('{')][0]; print('value', d['value'], 'ms', d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
timeout 300 python tools/timeline.py > gpurun_out/timeline_$TAG.txt 2>&1; tail -25 gpurun_out/timeline_$TAG.txt
