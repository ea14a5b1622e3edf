#!/bin/bash
# dense_tc development run: training goldens, training bench, profiler table, ncu captures of the two kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/dense_train_tests.log 2>&1; tail -n 3 gpurun_out/dense_train_tests.log
timeout 600 python bench.py --train > gpurun_out/dense_bench_train.json 2> gpurun_out/dense_bench_train.err; tail -c 1500 gpurun_out/dense_bench_train.json
timeout 300 python tools/profile_train.py 64 > /dev/null 2>&1; head -40 gpurun_out/train_profile.txt | cut -c1-100,170-260
for kn in lin_tc_kernel wgrad_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn --launch-skip 0 --launch-count 1 \
    -o gpurun_out/dense_$kn -f python tools/run_dense.py > gpurun_out/ncu_dense_$kn.log 2>&1
  ls -la gpurun_out/dense_$kn.ncu-rep
done
