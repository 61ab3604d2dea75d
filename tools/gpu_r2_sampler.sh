#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show() { python -c "import sys,json; d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('value', d['value'], 'ms', d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'loss', d['loss']['loss'])"; }
for f in 1 0 1; do echo "== LOFT_FUSED_SAMPLER=$f"; LOFT_FUSED_SAMPLER=$f timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_sampler$f.json | show; done
