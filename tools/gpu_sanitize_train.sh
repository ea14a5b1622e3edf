#!/bin/bash
# compute-sanitizer memcheck over the training-path kernels (tensor-core dense layers, fused cost volume, grouped rows) and a whole
# training step at a small size; racecheck over the shared-memory pipelines of the dense kernels
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_dense_tc.py tests/test_gpu_costvol_train.py \
  "tests/test_gpu_ops.py::test_query_and_group_rows_layout_equals_the_op_chain" "tests/test_gpu_train.py::test_train_step_is_bitwise_repeatable" \
  -m gpu -q -x -k "not long_reduction" > gpurun_out/memcheck_train.log 2>&1; echo "memcheck rc=$?" > gpurun_out/memcheck_train.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/run_dense.py 20000 256 256 > gpurun_out/racecheck_dense.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/memcheck_train.txt
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/memcheck_train.log gpurun_out/racecheck_dense.log >> gpurun_out/memcheck_train.txt
tail -n 3 gpurun_out/memcheck_train.log >> gpurun_out/memcheck_train.txt
cat gpurun_out/memcheck_train.txt
