#!/bin/bash
mkdir -p gpurun_out
AB_FLAGS="${AB_FLAGS:-59}" bash tools/gpu_ab.sh | tail -n 3
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1; tail -n 1 gpurun_out/timeline.log | cut -c1-400
