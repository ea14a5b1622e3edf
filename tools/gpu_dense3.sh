#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dev_dense_tc.py --big 2>&1 | grep -v "Warning\|Consider\|return float" > gpurun_out/dev_dense.log; tail -22 gpurun_out/dev_dense.log
timeout 900 python -m pytest tests/test_gpu_dense_tc.py tests/test_gpu_train.py "tests/test_gpu_ops.py::test_query_and_group_rows_layout_equals_the_op_chain" tests/test_gpu_backbone.py -x -q > gpurun_out/dense_tests.log 2>&1; tail -n 5 gpurun_out/dense_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lin_tc_kernel --launch-skip 0 --launch-count 1 -o gpurun_out/dense_lin_tc_kernel -f python tools/run_dense.py > gpurun_out/ncu_dense_lin.log 2>&1
ls -la gpurun_out/dense_lin_tc_kernel.ncu-rep
