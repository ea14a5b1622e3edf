#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -m gpu -q --tb=short -x 2>&1 | tail -n 6
for w in 1 0; do
  RT_BQ_WARP=$w python tools/stage_profile.py 32 10 > /dev/null 2>&1
  echo "RT_BQ_WARP=$w $(tail -1 gpurun_out/stage_profile.txt) | SA1 $(grep 'pn_head SA1' gpurun_out/stage_profile.txt | awk '{print $1}') SA2 $(grep 'pn_head SA2' gpurun_out/stage_profile.txt | awk '{print $1}') SA3 $(grep 'pn_head SA3' gpurun_out/stage_profile.txt | awk '{print $1}')"
  RT_BQ_WARP=$w python tools/bench_ops.py 2>/dev/null | grep ball_query
done | tee gpurun_out/ab_bq.txt
