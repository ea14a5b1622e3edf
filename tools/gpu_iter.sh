#!/bin/bash
# iteration loop of the inference work: GPU suite, A/B of the engine flags (tools/gpu_ab.sh), kernel timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v "^$" | tail -n 40 > gpurun_out/pytest_gpu.log
tail -n 14 gpurun_out/pytest_gpu.log
AB_FLAGS="${AB_FLAGS:-59}" bash tools/gpu_ab.sh | tail -n 2
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1; tail -n 2 gpurun_out/timeline.log
