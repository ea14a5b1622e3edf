#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense_tc.py tests/test_gpu_train.py -x -q > gpurun_out/dense_tests.log 2>&1; tail -n 3 gpurun_out/dense_tests.log | cut -c1-300
timeout 200 python tools/dev_dense_tc.py --big 2>&1 | grep "ms:" | cut -c1-140
timeout 300 python bench.py --train > gpurun_out/dense_bench_train.json 2> gpurun_out/dense_bench_train.err; cut -c1-260 gpurun_out/dense_bench_train.json
