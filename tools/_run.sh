timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short -x 2>&1 | tail -n 3
for g in 1 0; do RT_GRU_CLUSTER=$g timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cluster=$g B=32', round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms')"; done
for g in 1 0; do RT_GRU_CLUSTER=$g timeout 600 python bench.py --batch 1 --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cluster=$g B=1', round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms')"; done
python tools/stage_profile.py 32 10 | tail -4
