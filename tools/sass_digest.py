"""Opcode digest of the shipped library: `cuobjdump -sass` per kernel, counting the mnemonics that prove the Blackwell
path (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, SYNCS = mbarrier, REDUX, UCGABAR/cluster).
    python tools/sass_digest.py [ratrack_b200/libratrack_b200.so] > profiles/r2_sass_digest.txt
Runs anywhere cuobjdump is installed (no GPU needed)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "ratrack_b200/libratrack_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "REDUX", "UCGABAR", "MAPA", "FFMA2", "FADD2",
       "FMUL2", "LDG", "STG", "LDS", "STS", "SHFL", "HMMA", "BAR", "ATOMG", "RED", "NANOSLEEP"]
kern, arch = None, None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        counts[kern]["__arch__"] = arch
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["__total__"] += 1
        counts[kern][op] += 1
print(f"# cuobjdump -sass {so}: per-kernel SASS opcode counts (static instructions); arch of every cubin: "
      f"{sorted(set(c['__arch__'] for c in counts.values()))}")
print(f"# columns: total instructions | " + " ".join(KEY))
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in KEY))
for k, c in counts.items():
    print(f"{k[:58]:58s} {c['__total__']:6d} " + " ".join(f"{c[x]:7d}" if c[x] else f"{'.':>7s}" for x in KEY))
