#!/bin/bash
# round-2 evidence, second pass (one B200): full GPU tests, smoke, bench (+ train records), reference arm, per-op table, stage
# profile, training profile, ncu --set full captures (cost volume, tensor-core dense forward / wgrad), launch lists
mkdir -p gpurun_out
nproc > gpurun_out/host.log; nvidia-smi -L >> gpurun_out/host.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 8 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
timeout 600 python tools/bench_ops.py > gpurun_out/ops_roofline.txt 2> gpurun_out/ops_roofline.err
timeout 600 python tools/stage_profile.py 32 10 > /dev/null 2>&1
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 600 python tools/profile_train.py 256 > gpurun_out/train_profile.log 2>&1
timeout 300 python tools/dev_dense_tc.py --big 2>&1 | grep -v "Warning\|Consider\|return float" > gpurun_out/dense_tc_table.txt
bash tools/gpu_ncu_one.sh costvol_tc_kernel 0 r2_costvol_final
for kn in lin_tc_kernel wgrad_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn --launch-skip 0 --launch-count 1 \
    -o gpurun_out/r2_$kn -f python tools/run_dense.py > gpurun_out/ncu_r2_$kn.log 2>&1
done
bash tools/gpu_launches.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train.csv \
  python bench.py --train --batch 64 --steps 1 --warmup 1 > gpurun_out/launches_train.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_reference.json
