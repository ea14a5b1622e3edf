#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -k "fps or vod or reference_kernels" -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest_fps256.log
RT_FPS_THREADS=128 timeout 900 python -m pytest tests/test_gpu_ops.py -k "fps or vod" -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest_fps128.log
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_fused_v6.json 2> gpurun_out/bench_fused_v6.err
RT_FPS_THREADS=128 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_fused_v6_fps128.json 2>> gpurun_out/bench_fused_v6.err
tail -3 gpurun_out/pytest_fps256.log gpurun_out/pytest_fps128.log gpurun_out/pytest_gpu.log; tail -2 gpurun_out/timeline.log; cat gpurun_out/bench_fused_v6.json gpurun_out/bench_fused_v6_fps128.json | cut -c1-260
