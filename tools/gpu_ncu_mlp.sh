#!/bin/bash
# ncu --set full + source of ONE launch: the SA3 ns=32 gather kernel of pn_head (11th mlp_tc launch of a forward),
# and of the weighted_sum kernel; reports come back (small), CSV pages are cut here on the box too
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc --launch-skip 10 --launch-count 1 \
  -o gpurun_out/mlp_sa3 -f python tools/run_forward.py 32 1 > gpurun_out/ncu_mlp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:weighted_sum --launch-count 1 \
  -o gpurun_out/wsum -f python tools/run_forward.py 32 1 >> gpurun_out/ncu_mlp.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/ncu_mlp.log
