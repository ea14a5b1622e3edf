#!/bin/bash
# compute-sanitizer memcheck over the training-path kernels: one whole training step at batch 1 / N=256 with every dense layer
# forced onto the tensor-core kernels (dense_tc forward / dgrad / wgrad, absmax, group_rows, cv1, act_grad, wsum, segment sums),
# and the dense kernels alone at a cost-volume shape.  (The pytest files under memcheck take > 25 minutes: bounded runs instead.)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_train_small.py > gpurun_out/memcheck_train.log 2>&1; echo "train step rc=$?" > gpurun_out/memcheck_train.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_dense.py 20000 256 256 > gpurun_out/memcheck_dense.log 2>&1; echo "dense kernels rc=$?" >> gpurun_out/memcheck_train.txt
grep -h "ERROR SUMMARY" gpurun_out/memcheck_train.log gpurun_out/memcheck_dense.log >> gpurun_out/memcheck_train.txt
tail -n 1 gpurun_out/memcheck_train.log >> gpurun_out/memcheck_train.txt
cat gpurun_out/memcheck_train.txt
