#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 40 > gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_ops.py > gpurun_out/ops_roofline.txt 2>gpurun_out/ops_roofline.err; grep -E "knn" gpurun_out/ops_roofline.txt
AB_FLAGS="${AB_FLAGS:-59 123}" bash tools/gpu_ab.sh | tail -n 4
RT_ENGINE_FLAGS=123 timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1; tail -n 2 gpurun_out/timeline.log
bash tools/gpu_ncu_mlp.sh
