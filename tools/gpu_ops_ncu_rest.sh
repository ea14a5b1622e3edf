#!/bin/bash
# the Seam-A launches the 120 s window of tools/gpu_ops_ncu.sh did not reach (three_nn, three_interpolate, knn): same command,
# first 15 matching launches skipped, only the metrics of the TSV (a few passes per kernel instead of 39)
mkdir -p gpurun_out
M=gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
timeout 40 ncu --metrics $M --clock-control none \
  -k regex:'fps_|ball_query|three_nn|knn|gather_rows|scatter_rows|inverse_index|three_interpolate|segment_sum' --launch-skip 15 -c 14 \
  -o gpurun_out/r2_ops_rest -f python tools/bench_ops.py --once > gpurun_out/ops_ncu_rest.log 2>&1
tail -n 2 gpurun_out/ops_ncu_rest.log; ls -la gpurun_out/r2_ops_rest.ncu-rep
