#!/bin/bash
for t in 256 128; do RT_FPS_THREADS=$t python tools/bench_ops.py 2>/dev/null | grep furthest | sed "s/^/T=$t /"; done
for t in 256 128; do RT_FPS_THREADS=$t python tools/stage_profile.py 32 10 > /dev/null 2>&1; echo "T=$t $(cat gpurun_out/stage_profile.txt | awk '{printf "%s ", $1}')"; done
