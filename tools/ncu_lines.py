"""Warp-stall samples of a one-kernel ncu report aggregated per CUDA source line (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None
rows = []
hdr = None
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0]:   # a CUDA source line (SASS rows have an empty line number)
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        rows.append((int(r[i_s] or 0), int(r[i_e] or 0), cur_file, int(r[0]), r[1].strip()))
tot = sum(x[0] for x in rows)
print(f"# {rep}: {tot} samples; top {topn} source lines (samples, % , warp-instructions, file:line, text)")
for s, e, f, ln, txt in sorted(rows, key=lambda x: -x[0])[:topn]:
    print(f"{s:7d} {100.0 * s / max(tot, 1):5.1f}% {e:10d}  {f}:{ln:<4d} {txt[:110]}")
