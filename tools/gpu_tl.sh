#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short 2>&1 | tail -n 6 > gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1; tail -n 4 gpurun_out/timeline.log
for rep in 1 2; do timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']), ' costvol ms', round(d['roofline']['avg_launch_ms'],3))"; done
