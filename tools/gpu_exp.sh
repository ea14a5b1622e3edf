#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -m gpu -q --tb=short -x -k "fps or backbone or vod or varlen or reference_kernels" 2>&1 | tail -n 3
python tools/bench_ops.py 2>/dev/null | grep furthest
for round in 1 2; do
python tools/stage_profile.py 32 10 > /dev/null 2>&1; cat gpurun_out/stage_profile.txt | awk '{printf "%s ", $1}'; echo
done
