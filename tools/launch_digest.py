"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) per kernel:
    python tools/launch_digest.py launches.csv FORWARDS > profiles/rN_launches.txt
Per-launch times under ncu are cold-cache and serialised: only each kernel's SHARE of the step is comparable with the
CUDA-event timings of bench.py."""
import collections
import csv
import re
import sys

path, fwd = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[0] == "ID" or len(r) <= iv:
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    us = v / 1e3 if r[iu].startswith("ns") else (v * 1e3 if r[iu].startswith("ms") else v)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", r[ik])
    name = name.split("(")[0]
    own = not (name.startswith("at::") or "cub::" in name or "elementwise" in name or name.startswith("std::"))
    a = agg.setdefault(name, [0, 0.0, own])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values() if a[2]) / fwd
n = sum(a[0] for a in agg.values() if a[2]) / fwd
print(f"# own kernels only: {tot:.0f} us per forward, {n:.0f} launches per forward ({fwd} forwards aggregated)")
for name, (cnt, us, own) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if own:
        print(f"{us / fwd:8.0f} us/fwd {100 * us / fwd / tot:6.1f}%  x{cnt / fwd:4.0f}  avg {us / cnt:8.1f} us  {name}")
other = [(n_, a) for n_, a in agg.items() if not a[2]]
if other:
    print("# torch kernels in the same capture (input staging of the driver script): " +
          ", ".join(f"{n_} x{a[0]}" for n_, a in other))
