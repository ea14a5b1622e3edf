#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/parity_report.py > gpurun_out/parity.log 2>&1
timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_ops.py -k "backbone or knn_expanded or tensor_core or odd" -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fused_v2.json 2> gpurun_out/bench_fused_v2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -16 gpurun_out/parity.log; tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_fused_v2.json; tail -5 gpurun_out/bench_fused_v2.err
