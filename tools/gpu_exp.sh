#!/bin/bash
for d in 64 16 32 48 63; do
  RT_CV_DEBUG=$d python tools/stage_profile.py 32 5 > /dev/null 2>&1
  echo "debug $d: costvol $(grep 'cost volume' gpurun_out/stage_profile.txt | awk '{print $1}') us"
done
