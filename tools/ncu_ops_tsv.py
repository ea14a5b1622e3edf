"""Raw-page extract of a multi-kernel ncu report as a TSV (one row per captured launch; units as ncu prints them).
    python tools/ncu_ops_tsv.py gpurun_out/ops_full.ncu-rep > profiles/rN_ncu_ops_full.tsv"""
import csv
import subprocess
import sys

COLS = [("Kernel Name", "Kernel Name"), ("gpu__time_duration.sum", "gpu__time_duration"), ("launch__grid_size", "launch__grid_size"),
        ("launch__block_size", "launch__block_size"), ("launch__registers_per_thread", "launch__registers_per_thread"),
        ("dram__bytes_read.sum", "dram__bytes_read"), ("dram__bytes_write.sum", "dram__bytes_write"),
        ("lts__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.%elapsed"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.%elapsed"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.%elapsed"),
        ("sm__inst_executed.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.%elapsed"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.%active"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.%active"),
        ("smsp__inst_executed.sum", "smsp__inst_executed")]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
have = [(k, label) for k, label in COLS if k in h]
print("\t".join(label + (f" [{units[h.index(k)]}]" if units[h.index(k)] else "") for k, label in have))
for r in rows[2:]:
    if len(r) != len(h):
        continue
    print("\t".join((r[h.index(k)][:72] if k == "Kernel Name" else r[h.index(k)]) for k, _ in have))
