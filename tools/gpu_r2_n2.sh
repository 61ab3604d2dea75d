#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 60 --warmup 5 --no-cpu-baseline; }
show() { python -c "import sys,json; d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print(d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value'])"; }
echo "== gc managed, reserve 8"; run 29511 2>gpurun_out/n2_a.err | show
echo "== gc auto, reserve 8"; LOFT_GC=auto run 29512 2>gpurun_out/n2_b.err | show
echo "== gc managed, no overlap"; LOFT_OVERLAP_COMM=0 run 29513 2>gpurun_out/n2_c.err | show
