#!/bin/bash
# hardware probe of the TMA gather4 instruction + A/B of the wide (16 worker warps) mlp_tc variant (RT_MLP_WIDE)
timeout 120 tools/tma_gather_probe
echo "--- wide mlp_tc A/B"
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short -x 2>&1 | tail -n 2
for wv in 0 1; do RT_MLP_WIDE=$wv python tools/stage_profile.py 32 10 > /dev/null 2>&1; echo "RT_MLP_WIDE=$wv $(cat gpurun_out/stage_profile.txt | awk '{printf "%s ", $1}')"; done
