#!/bin/bash
# compute-sanitizer memcheck over a small fused forward, the Seam-A op tests (subset) and the association kernel
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_forward.py 2 1 > gpurun_out/memcheck_forward.log 2>&1; echo "forward rc=$?" > gpurun_out/memcheck.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py tests/test_association.py -m gpu -q -x -k "fps_bit_exact or ball_query_bit_exact or three_nn or knn_bit_exact or interpolate or group_and_gather or sinkhorn" > gpurun_out/memcheck_ops.log 2>&1; echo "ops rc=$?" >> gpurun_out/memcheck.txt
grep -h "ERROR SUMMARY" gpurun_out/memcheck_forward.log gpurun_out/memcheck_ops.log >> gpurun_out/memcheck.txt
tail -n 3 gpurun_out/memcheck_ops.log >> gpurun_out/memcheck.txt
cat gpurun_out/memcheck.txt
