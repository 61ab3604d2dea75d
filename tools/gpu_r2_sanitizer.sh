#!/bin/bash
# compute-sanitizer over the kernel parity tests: memcheck (all kernels, incl. the cta_group::2 pair
# form and the polygon rasteriser) and racecheck (shared-memory hazards of the non-GEMM kernels; the
# GEMM's shared memory is written by TMA / read by tcgen05, which racecheck does not model).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== memcheck: tests/test_gpu_kernels.py tests/test_pipeline.py tests/test_gpu_proposals.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_pipeline.py tests/test_gpu_proposals.py -m gpu -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -20
echo "rc=$?"
echo "== racecheck: non-GEMM kernels"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_pipeline.py -m gpu -x -q -k "roi_align or mask_target or iou_assign or nms or coders or rot90 or elem_losses or focal or sgd or polygon or bonai_dataset" > gpurun_out/r02_racecheck_full.log 2>&1
echo "rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r02_racecheck_full.log | head
grep -E "hazard detected|Error:|Warning:" gpurun_out/r02_racecheck_full.log | sed 's/(threadIdx.*//' | sort | uniq -c | sort -rn | head -20
} 2>&1 | tee gpurun_out/r02_sanitizer.txt
