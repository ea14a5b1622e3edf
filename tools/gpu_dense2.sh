#!/bin/bash
# training-path development run: parity tests of the new kernels, training goldens, training bench, profiler table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense_tc.py tests/test_gpu_train.py "tests/test_gpu_ops.py::test_query_and_group_rows_layout_equals_the_op_chain" tests/test_gpu_backbone.py -x -q > gpurun_out/dense_tests.log 2>&1; tail -n 5 gpurun_out/dense_tests.log
timeout 600 python bench.py --train > gpurun_out/dense_bench_train.json 2> gpurun_out/dense_bench_train.err; cut -c1-400 gpurun_out/dense_bench_train.json
timeout 300 python tools/profile_train.py 64 > /dev/null 2>&1; head -45 gpurun_out/train_profile.txt | cut -c1-100,170-260
