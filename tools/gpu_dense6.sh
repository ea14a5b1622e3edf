#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_costvol_train.py tests/test_gpu_dense_tc.py tests/test_gpu_parity_floor.py -x -q > gpurun_out/dense_tests.log 2>&1; grep -v "^$" gpurun_out/dense_tests.log | tail -n 6 | cut -c1-300
timeout 600 python bench.py --train > gpurun_out/dense_bench_train.json 2> gpurun_out/dense_bench_train.err; cut -c1-300 gpurun_out/dense_bench_train.json; grep -o '"peak_mem_gb": [0-9.]*' gpurun_out/dense_bench_train.json
bash tools/gpu_sanitize_train.sh
