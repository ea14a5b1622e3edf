#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_ops.py -k "backbone or knn_expanded or tensor_core or odd" -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_fused_v3.json 2> gpurun_out/bench_fused_v3.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --batch 128 > gpurun_out/bench_fused_v3_b128.json 2>> gpurun_out/bench_fused_v3.err
tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_fused_v3.json gpurun_out/bench_fused_v3_b128.json; tail -5 gpurun_out/bench_fused_v3.err
