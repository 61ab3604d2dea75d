#!/bin/bash
# one GPU visit: kernel parity tests + whole-step parity
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python tools/e2e_check.py 256 1 10 > gpurun_out/e2e_256.log 2>&1
echo "e2e rc=$?" >> gpurun_out/e2e_256.log
tail -40 gpurun_out/e2e_256.log
timeout 1200 python tools/e2e_check.py 1024 2 80 > gpurun_out/e2e_1024.log 2>&1
echo "e2e1024 rc=$?" >> gpurun_out/e2e_1024.log
grep -E "worst|rc=|Error|error" gpurun_out/e2e_1024.log | head
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
tail -5 gpurun_out/bench.log
