#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short 2>&1 | tail -n 8 > gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_fused_v7.json 2> gpurun_out/bench_fused_v7.err
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/timeline.log; cut -c1-330 gpurun_out/bench_fused_v7.json; grep -o '"roofline".*' gpurun_out/bench_fused_v7.json | cut -c1-400
