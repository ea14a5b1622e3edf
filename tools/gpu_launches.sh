#!/bin/bash
# ncu launch list of a short bench run: 64 pairs = 2 micro-batches of 32 per step
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --total 64 --steps 2 --warmup 1 --no-cpu --no-train > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log | cut -c1-300
