#!/bin/bash
cd "$(dirname "$0")/.."
show() { python -c "import sys,json; d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('value', d['value'], 'ms', d['ms_per_step'], d['step_ms']['p50'], 'e2e', d['e2e']['value'])"; }
for v in step_end forward; do echo "== LOFT_PREFETCH_AT=$v"; LOFT_PREFETCH_AT=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | show; done
echo "== LOFT_PREFETCH=0"; LOFT_PREFETCH=0 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | show
