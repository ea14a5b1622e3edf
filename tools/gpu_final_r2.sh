#!/bin/bash
# round-2 evidence (one B200): full GPU tests, smoke, parity floor, bench (+ CPU baseline + ref_gpu + train records), reference
# arm, per-op table, stage profile, kernel timeline, ncu --set full captures of the top kernels, ncu launch list of a short bench
mkdir -p gpurun_out
nproc > gpurun_out/host.log; nvidia-smi -L >> gpurun_out/host.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 8 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
timeout 600 python tools/bench_ops.py > gpurun_out/ops_roofline.txt 2> gpurun_out/ops_roofline.err
timeout 600 python tools/stage_profile.py 32 10 > /dev/null 2>&1
timeout 600 python tools/timeline.py 32 > gpurun_out/timeline.log 2>&1
timeout 600 python tools/profile_train.py > gpurun_out/train_profile.txt 2>&1
bash tools/gpu_ncu_one.sh costvol_tc_kernel 0 r2_costvol_final
bash tools/gpu_ncu_one.sh sa_tc_kernel 5 r2_sa3_final
bash tools/gpu_ncu_one.sh ball_query_thread_kernel 0 r2_bq_final
bash tools/gpu_launches.sh
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_reference.json
