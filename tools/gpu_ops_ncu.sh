#!/bin/bash
# ncu --set full of every Seam-A kernel (one launch each), exported to CSV on the box (the .ncu-rep is too big to bring back)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'fps_reg|ball_query|three_nn|knn|gather|group|three_interpolate' \
  -o /tmp/ops_full -f python tools/bench_ops.py --once > gpurun_out/ops_ncu.log 2>&1
ncu -i /tmp/ops_full.ncu-rep --page raw --csv > gpurun_out/ops_full_raw.csv 2>/dev/null
ls -la /tmp/ops_full.ncu-rep gpurun_out/ops_full_raw.csv
