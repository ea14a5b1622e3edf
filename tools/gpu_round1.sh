#!/bin/bash
# first GPU pass: parity tests, reference-kernel golden, smoke, modular bench, launch list
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python -m oracle.gen_golden_ref_gpu > gpurun_out/gen_ref.log 2>&1
timeout 300 python tools/probe_sqdist.py > gpurun_out/probe.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_modular.json 2> gpurun_out/bench_modular.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_modular.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_modular.json
