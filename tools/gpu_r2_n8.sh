#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline; }
show() { python -c "import sys,json; d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print(d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value'])"; }
echo "== N=8 overlap on"; LOFT_TIME_COMM=1 run 29521 2>gpurun_out/n8_overlap.err | tee gpurun_out/n8_overlap.json | show; grep "exposed" gpurun_out/n8_overlap.err | head -3
echo "== N=8 overlap off"; LOFT_OVERLAP_COMM=0 LOFT_TIME_COMM=1 run 29522 2>gpurun_out/n8_nooverlap.err | tee gpurun_out/n8_nooverlap.json | show; grep "exposed" gpurun_out/n8_nooverlap.err | head -3
