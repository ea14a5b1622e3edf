#!/bin/bash
# last check of the round: the whole GPU suite (per-test timeout from pytest.ini), smoke, association latency
mkdir -p gpurun_out
nproc > gpurun_out/host.log; nvidia-smi -L >> gpurun_out/host.log
timeout 700 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 8 > gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 120 python tools/bench_track.py 1024 > gpurun_out/track_bench_1024.txt 2>&1; timeout 120 python tools/bench_track.py 320 > gpurun_out/track_bench_320.txt 2>&1
cat gpurun_out/track_bench_1024.txt gpurun_out/track_bench_320.txt > gpurun_out/track_bench.txt
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; cat gpurun_out/track_bench.txt
