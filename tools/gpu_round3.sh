#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/parity_report.py > gpurun_out/parity.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fused_v1.json 2> gpurun_out/bench_fused_v1.err
tail -12 gpurun_out/parity.log; cat gpurun_out/bench_fused_v1.json; tail -5 gpurun_out/bench_fused_v1.err
