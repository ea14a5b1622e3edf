#!/bin/bash
# round-2 captures: ncu --set full of the cost volume and of the SA3 ns=32 scale (sa_tc), A/B of the sa_tc flag
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short -x 2>&1 | tail -n 5
bash tools/gpu_ncu_one.sh costvol_tc_kernel 0 r2_costvol_v3
bash tools/gpu_ncu_one.sh sa_tc_kernel 3 r2_sa3_ns32
for round in 1 2; do
  for f in 571 1595; do
    RT_ENGINE_FLAGS=$f python tools/stage_profile.py 32 10 > /dev/null 2>&1
    echo "round $round flags=$f $(tail -1 gpurun_out/stage_profile.txt) | $(grep -E 'pn_head SA2|pn_head SA3|mse SA2|mse SA3' gpurun_out/stage_profile.txt | awk '{printf "%s ", $1}')"
  done
done | tee gpurun_out/ab_sa_tc.txt
