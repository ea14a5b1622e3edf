"""Digest of one-kernel ncu reports (raw + source pages): headline metrics, stall mix, hottest SASS lines, opcode mix.
    python tools/ncu_digest.py gpurun_out/x.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, r = rows[0], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    if w in h:
        print(f"{w:86s} {r[h.index(w)]}")
print("-- stall reasons (warps per issue-active cycle)")
for k, v in zip(h, r):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            if float(v) > 0.15:
                print(f"   {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {float(v):.2f}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hh = rows[1]
ix = {k: i for i, k in enumerate(hh)}
data = [x for x in rows[2:] if len(x) == len(hh)]
tot = sum(int(x[ix["# Samples"]] or 0) for x in data)
print("-- hottest SASS lines (samples of", tot, "| executions | smem wavefronts | global tag requests)")
for x in sorted(data, key=lambda x: -int(x[ix["# Samples"]] or 0))[:topn]:
    print(x[ix["# Samples"]].rjust(6), x[ix["Instructions Executed"]].rjust(9), (x[ix["L1 Wavefronts Shared"]] or "").rjust(9),
          (x[ix["L1 Tag Requests Global"]] or "").rjust(9), x[ix["Source"]][:96])
c, w, g, smp = (collections.Counter() for _ in range(4))
for x in data:
    s = x[ix["Source"]].split()
    if not s:
        continue
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    c[op] += int(x[ix["Instructions Executed"]] or 0)
    w[op] += int(x[ix["L1 Wavefronts Shared"]] or 0)
    g[op] += int(x[ix["L1 Tag Requests Global"]] or 0)
    smp[op] += int(x[ix["# Samples"]] or 0)
print("-- instructions by opcode:", c.most_common(14))
print("-- smem wavefronts:", w.most_common(4), "| global tag requests:", g.most_common(4))
print("-- samples by opcode:", smp.most_common(10))
