#!/bin/bash
# association path: its GPU tests, whole-forward latency at N=1024 / N=320 (tools/bench_track.py), memcheck of the training kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_track.py tests/test_association.py -q -x > gpurun_out/track_tests.log 2>&1; tail -n 4 gpurun_out/track_tests.log | cut -c1-300
timeout 300 python tools/bench_track.py 1024 > gpurun_out/track_bench_1024.txt 2>&1; timeout 300 python tools/bench_track.py 320 > gpurun_out/track_bench_320.txt 2>&1
cat gpurun_out/track_bench_1024.txt gpurun_out/track_bench_320.txt > gpurun_out/track_bench.txt; cat gpurun_out/track_bench.txt
bash tools/gpu_sanitize_train.sh
