#!/bin/bash
# ncu --set full + source of ONE launch: $1 = kernel regex, $2 = launches to skip, $3 = output name
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 --launch-count 1 \
  -o gpurun_out/$3 -f python tools/run_forward.py 32 1 > gpurun_out/ncu_$3.log 2>&1
ls -la gpurun_out/$3.ncu-rep; tail -n 1 gpurun_out/ncu_$3.log
