#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_fused_v5.json 2> gpurun_out/bench_fused_v5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:costvol_tc -s 1 -c 1 -f -o gpurun_out/prof_costvol python tools/run_forward.py 32 2 > gpurun_out/ncu_costvol.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 44 -c 1 -f -o gpurun_out/prof_mlp_sa3 python tools/run_forward.py 32 2 > gpurun_out/ncu_mlp.log 2>&1
cat gpurun_out/bench_fused_v5.json; tail -3 gpurun_out/ncu_costvol.log gpurun_out/ncu_mlp.log; ls -la gpurun_out/*.ncu-rep
