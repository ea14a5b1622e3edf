#!/bin/bash
# end-of-milestone evidence: full GPU tests, smoke, parity report, bench (with CPU baseline + ref_gpu), launch list, timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 8 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python tools/parity_report.py > gpurun_out/parity.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_fused.csv python tools/run_forward.py 32 3 > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; cat gpurun_out/bench_final.json; cat gpurun_out/bench_reference.json | cut -c1-200
