#!/bin/bash
# restored-checkpoint sanity: full GPU test-suite, smoke, one bench line
mkdir -p gpurun_out
nproc > gpurun_out/host.log; nvidia-smi -L >> gpurun_out/host.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 12 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --no-cpu > gpurun_out/bench_sanity.json 2> gpurun_out/bench_sanity.err
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_sanity.json
