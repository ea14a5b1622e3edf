#!/bin/bash
# copy the evidence of tools/gpu_final.sh from gpurun_out/ (scratch) into profiles/ (tracked); run in the dev container
set -e
cd "$(dirname "$0")/.."
R=${1:-r1}
cp gpurun_out/bench_final.json profiles/${R}_bench_final.json
cp gpurun_out/bench_reference.json profiles/${R}_bench_reference_arm.json
cp gpurun_out/bench_train.json profiles/${R}_bench_train.json
cp gpurun_out/ops_roofline.json profiles/${R}_ops_roofline.json
cp gpurun_out/ops_roofline.txt profiles/${R}_ops_roofline.txt
cp gpurun_out/timeline.txt profiles/${R}_timeline_final.txt
{ echo "# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised) of: python bench.py --steps 2 --warmup 1 --no-cpu"
  echo "# = 3 device-resident forwards + 4 forwards of the end-to-end leg; one-time kernels (pack_umma) belong to engine creation"
  python tools/launch_digest.py gpurun_out/launches_bench.csv 7; } > profiles/${R}_launches_fused_final.txt
{ cat gpurun_out/host.log; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; } > profiles/${R}_gpu_tests.txt
[ -f gpurun_out/parity_report.txt ] && cp gpurun_out/parity_report.txt profiles/${R}_parity_report.txt
ls profiles
