#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/tc_probe > gpurun_out/tc_probe.log 2>&1
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short 2>&1 | tail -120 > gpurun_out/pytest_backbone.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fused_v0.json 2> gpurun_out/bench_fused_v0.err
cat gpurun_out/tc_probe.log; tail -15 gpurun_out/pytest_backbone.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_fused_v0.json; tail -5 gpurun_out/bench_fused_v0.err
