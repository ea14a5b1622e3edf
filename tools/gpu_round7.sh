#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_fused_v4.json 2> gpurun_out/bench_fused_v4.err
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/timeline.log; cat gpurun_out/bench_fused_v4.json; tail -5 gpurun_out/bench_fused_v4.err
