#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_track.py -m gpu -q --tb=short -x 2>&1 | tail -n 4
for round in 1 2; do
python tools/stage_profile.py 32 10 > /dev/null 2>&1; cat gpurun_out/stage_profile.txt | awk '{printf "%s ", $1}'; echo
done
