#!/bin/bash
# iteration loop of the training-path work: parity tests of the touched kernels, bench.py --train, the profiler rows of the cost-volume kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_costvol_train.py tests/test_gpu_train.py "tests/test_gpu_ops.py::test_query_and_group_rows_layout_equals_the_op_chain" -x -q > gpurun_out/dense_tests.log 2>&1; tail -n 3 gpurun_out/dense_tests.log | cut -c1-300
timeout 300 python bench.py --train > gpurun_out/dense_bench_train.json 2> gpurun_out/dense_bench_train.err; cut -c1-260 gpurun_out/dense_bench_train.json; grep -o '"peak_mem_gb": [0-9.]*' gpurun_out/dense_bench_train.json
timeout 200 python tools/profile_train.py 64 > /dev/null 2>&1; grep "wsum\|cv1\|act_grad\|GroupRows\|group_rows" gpurun_out/train_profile.txt | cut -c1-90,170-260
