#!/bin/bash
# A/B of engine scheduling flags (RT_ENGINE_FLAGS bits: 1 costvol_tc, 2 mlp_tc, 4 two lanes, 8 FPS SM-exclusive,
# 16 early kNN, 32 own prioritised main stream, 64 Morton-order cost-volume tiles, 128 two clouds per FPS CTA)
# + parity of the backbone tests under TEST_FLAGS (default: the library default)
mkdir -p gpurun_out
RT_ENGINE_FLAGS=${TEST_FLAGS:-} timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short 2>&1 | tail -n 6 > gpurun_out/pytest_backbone.log
[ -z "${TEST_FLAGS:-}" ] && unset RT_ENGINE_FLAGS
tail -n 3 gpurun_out/pytest_backbone.log
line() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']), ' costvol ms', round(d['roofline']['avg_launch_ms'],3))"; }
for f in ${AB_FLAGS:-59 187}; do
  for rep in 1 2; do
    RT_ENGINE_FLAGS=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | line "flags=$f rep$rep"
  done
done | tee gpurun_out/ab.log
