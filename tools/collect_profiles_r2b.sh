#!/bin/bash
# copy the evidence of tools/gpu_final_r2b.sh (+ gpu_sanitize_train.sh, gpu_multi.sh) from gpurun_out/ (scratch) into profiles/ (tracked)
set -e
cd "$(dirname "$0")/.."
R=r2
cp gpurun_out/bench_final.json profiles/${R}_bench_final.json
cp gpurun_out/bench_reference.json profiles/${R}_bench_reference_arm.json
cp gpurun_out/ops_roofline.json profiles/${R}_ops_roofline.json
cp gpurun_out/ops_roofline.txt profiles/${R}_ops_roofline.txt
cp gpurun_out/timeline.txt profiles/${R}_timeline.txt
cp gpurun_out/stage_profile.txt profiles/${R}_stage_profile.txt
cp gpurun_out/train_profile.txt profiles/${R}_train_profile.txt
cp gpurun_out/dense_tc_table.txt profiles/${R}_dense_tc_table.txt
[ -f gpurun_out/track_bench.txt ] && cp gpurun_out/track_bench.txt profiles/${R}_track_bench.txt
[ -f gpurun_out/memcheck_train.txt ] && cp gpurun_out/memcheck_train.txt profiles/${R}_memcheck_train.txt
{ echo "# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised) of: python bench.py --total 64 --steps 2 --warmup 1 --no-cpu --no-train"
  echo "# = 6 device-resident forwards + 6 forwards of the end-to-end leg (2 micro-batches of 32 pairs per step); one-time kernels (pack_umma) belong to engine creation"
  python tools/launch_digest.py gpurun_out/launches_bench.csv 12; } > profiles/${R}_launches.txt
{ echo "# ncu launch list of: python bench.py --train --batch 64 --steps 1 --warmup 1  (two training steps: forward + loss + backward + Adam, batch 64, N=1024)"
  echo "# torch / cuDNN / cuBLAS kernels of the step are listed with the package's own (at::* element-wise kernels are summarised on the last line)"
  python tools/launch_digest.py gpurun_out/launches_train.csv 2 | cut -c1-200; } > profiles/${R}_launches_train.txt
{ cat gpurun_out/host.log; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log; } > profiles/${R}_gpu_tests.txt
{ echo "# ncu --set full --clock-control none --import-source on, one launch, B=32 N=1024 (python tools/run_forward.py 32 1); digest by tools/ncu_digest.py + tools/ncu_lines.py"
  python tools/ncu_digest.py gpurun_out/r2_costvol_final.ncu-rep 24; python tools/ncu_lines.py gpurun_out/r2_costvol_final.ncu-rep 24; } > profiles/${R}_ncu_costvol_tc.txt
{ echo "# ncu --set full, one launch: lin_tc_kernel forward, 1 M rows x 256 -> 256 (python tools/run_dense.py): the cost-volume layer shape of the training step"
  python tools/ncu_digest.py gpurun_out/r2_lin_tc_kernel.ncu-rep 16; } > profiles/${R}_ncu_lin_tc.txt
{ echo "# ncu --set full, one launch: wgrad_tc_kernel, dW (256 x 256) = dY^T . X over 1 M rows (python tools/run_dense.py)"
  python tools/ncu_digest.py gpurun_out/r2_wgrad_tc_kernel.ncu-rep 16; } > profiles/${R}_ncu_wgrad_tc.txt
python - <<PY
import csv, json, subprocess
raw = subprocess.run(["ncu", "-i", "gpurun_out/r2_costvol_final.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, r = rows[0], rows[1], rows[2]
def val(k):
    i = h.index(k); v = float(r[i].replace(",", "")); unit = u[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
json.dump({"kernel": "costvol_tc_kernel", "batch": 32, "points": 1024, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
           "source": "profiles/r2_ncu_costvol_tc.txt (ncu --set full, one launch, B=32 N=1024, the round-2 kernel)"},
          open("profiles/costvol_traffic.json", "w"), indent=1)
print("costvol traffic", rd + wr)
PY
ls profiles | wc -l
