#!/bin/bash
# iteration check: GPU tests (ops + backbone + train), per-op table, bench with one and two lanes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -n 15 > gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_ops.py > gpurun_out/ops_roofline.txt 2>gpurun_out/ops_roofline.err; grep -E "three_nn|ball_query|knn" gpurun_out/ops_roofline.txt
AB_FLAGS="${AB_FLAGS:-59 63}" bash tools/gpu_ab.sh | tail -n 4
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1; tail -n 2 gpurun_out/timeline.log
