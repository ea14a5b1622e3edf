#!/bin/bash
# one ncu --set full capture per Seam-A kernel of the current library (tools/bench_ops.py --once), then the GPU suite
mkdir -p gpurun_out
timeout 120 ncu --set full --clock-control none --import-source on \
  -k regex:'fps_|ball_query|three_nn|knn|gather_rows|scatter_rows|inverse_index|three_interpolate|segment_sum' -c 40 \
  -o gpurun_out/r2_ops_full -f python tools/bench_ops.py --once > gpurun_out/ops_ncu.log 2>&1
tail -n 3 gpurun_out/ops_ncu.log; ls -la gpurun_out/r2_ops_full.ncu-rep
timeout 90 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 3 | tee gpurun_out/pytest_gpu_last.log
