#!/bin/bash
# A/B of engine flags on the same box: 7 = all on, 3 = single lane
mkdir -p gpurun_out
for rep in 1 2; do
for f in 7 3; do
RT_ENGINE_FLAGS=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flags $f rep $rep', round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']), ' costvol ms', round(d['roofline']['avg_launch_ms'],3))"
done; done
RT_ENGINE_FLAGS=7 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --batch 128 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B128 flags 7', round(d['value']), 'frames/s', round(d['ms_per_step'],3))"
RT_ENGINE_FLAGS=3 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --batch 128 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B128 flags 3', round(d['value']), 'frames/s', round(d['ms_per_step'],3))"
