#!/bin/bash
# end-of-milestone evidence (one B200): full GPU tests, smoke, parity report, bench (+ CPU baseline + ref_gpu),
# reference arm, per-op roofline table, training-step bench, ncu launch list of the bench command, kernel timeline
mkdir -p gpurun_out
nproc > gpurun_out/host.log; nvidia-smi -L >> gpurun_out/host.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 8 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python tools/parity_report.py > gpurun_out/parity.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
timeout 600 python tools/bench_ops.py > gpurun_out/ops_roofline.txt 2> gpurun_out/ops_roofline.err
timeout 900 python bench.py --train --steps 5 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; cat gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_reference.json
cat gpurun_out/bench_train.json; tail -n 2 gpurun_out/bench_train.err
