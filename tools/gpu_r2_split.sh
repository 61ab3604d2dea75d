#!/bin/bash
# gradient exchange in three parts (heads under the trunk backward, layer3/4+FPN+RPN under layer2's
# backward, the rest after): A/B at N GPUs.   usage: gpu_r2_split.sh N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline; }
show() { python -c "import sys,json; d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print(d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value'], 'loss', d['loss']['loss'])"; }
for s in ${ORDER:-1 0 1 0}; do
  echo "== N=$N LOFT_SPLIT_COMM=$s"; LOFT_SPLIT_COMM=$s LOFT_TIME_COMM=1 run $((29530 + s)) 2>gpurun_out/n${N}_split$s.err | tee gpurun_out/n${N}_split$s.json | show
  grep "exposed" gpurun_out/n${N}_split$s.err | head -2; grep -i "error\|Traceback" gpurun_out/n${N}_split$s.err | head -3
done
