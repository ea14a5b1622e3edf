#!/bin/bash
# per-op roofline table (ours vs the reference's kernels) + one ncu --set full capture of every Seam-A kernel
mkdir -p gpurun_out
timeout 600 python tools/bench_ops.py > gpurun_out/ops_roofline.txt 2> gpurun_out/ops_roofline.err
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'fps_reg|ball_query|three_nn|knn|gather|group|three_interpolate' \
  -o gpurun_out/ops_full -f python tools/bench_ops.py --once > gpurun_out/ops_ncu.log 2>&1
cat gpurun_out/ops_roofline.txt; tail -n 3 gpurun_out/ops_roofline.err; tail -n 3 gpurun_out/ops_ncu.log; ls -la gpurun_out/ops_full.ncu-rep
