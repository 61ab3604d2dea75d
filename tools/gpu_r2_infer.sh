#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_infer.py -m gpu -x -q -k "soft_nms or inference or infer" 2>&1 | tail -3
for f in 1 0; do echo "== LOFT_SOFT_NMS_FAST=$f"; LOFT_SOFT_NMS_FAST=$f timeout 600 python tools/infer_bench.py --batches 1 4 16 32 --max-dets 100 2>/dev/null | tee gpurun_out/r02_infer_sweep_fast$f.jsonl | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['batch'], d['value'], d['ms_per_img'], d['breakdown_ms_per_img'])"; done
