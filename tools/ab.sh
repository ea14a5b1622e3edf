#!/bin/bash
# A/B timing on ONE box: alternate environment settings, several rounds each (tools/stage_profile.py totals).
#   tools/ab.sh "RT_MLP_PDL=0" "RT_MLP_PDL=1" ...
for round in 1 2 3; do
  for cfg in "$@"; do
    env $cfg python tools/stage_profile.py 32 10 > /dev/null 2>&1
    echo "round $round [$cfg] $(tail -1 gpurun_out/stage_profile.txt) | costvol $(grep 'cost volume' gpurun_out/stage_profile.txt | awk '{print $1}')"
  done
done
