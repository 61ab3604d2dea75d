#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/pipeline_bench.py 2>&1 | tail -3 | tee gpurun_out/r02_pipeline_bench.txt
