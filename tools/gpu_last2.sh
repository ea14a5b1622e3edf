#!/bin/bash
# closing run of the round on the committed state: the whole GPU suite, smoke, the bench line
mkdir -p gpurun_out
nproc > gpurun_out/host.log; nvidia-smi -L >> gpurun_out/host.log
timeout 500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 4 > gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -n 2 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_final.json
