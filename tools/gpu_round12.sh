#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 12 > gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_fused_v8.json 2> gpurun_out/bench_fused_v8.err
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/timeline.log; cut -c1-330 gpurun_out/bench_fused_v8.json
