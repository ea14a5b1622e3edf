"""Kernel timeline of one fused forward (torch.profiler / CUPTI): name, stream, start, duration.
Writes gpurun_out/timeline.txt.  Shows cross-stream overlap, which the serialised ncu launch list cannot."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200 import synthetic  # noqa: E402
from ratrack_b200.model_utils import Track4DBackbone  # noqa: E402


class Args:
    npoints = 512


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = Track4DBackbone(Args())
net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
net = net.cuda().eval()
d = synthetic.make_batch(B, 1024, seed=1234)
t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
h = torch.zeros(5, B, 128, device="cuda")
with torch.no_grad():
    for _ in range(3):
        net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h)
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
# stream ids come from the Chrome trace (FunctionEvent does not carry them)
streams = {}
try:
    import json
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        prof.export_chrome_trace(f.name)
        tr = json.load(open(f.name))
    for ev in tr.get("traceEvents", []):
        if ev.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "ts" in ev:
            streams[(ev["name"][:40], round(float(ev["dur"]), 1))] = ev.get("args", {}).get("stream", ev.get("tid"))
except Exception as ex:  # informational only
    print("no stream ids:", ex)
lines = []
busy_until = 0.0
idle = 0.0
tot = {}
for e in evs:
    st, en = e.time_range.start - t0, e.time_range.end - t0
    if st > busy_until:
        idle += st - busy_until
    busy_until = max(busy_until, en)
    sid = streams.get((e.name[:40], round(float(en - st), 1)), "?")
    short = e.name.replace("(anonymous namespace)::", "").replace("void ", "")[:64]
    lines.append(f"{st:9.1f} +{en - st:8.1f} us  s{sid}  {short}")
    key = short.split("(")[0]
    tot[key] = tot.get(key, (0, 0.0))
    tot[key] = (tot[key][0] + 1, tot[key][1] + en - st)
end = max(e.time_range.end for e in evs)
lines.append(f"span {end - t0:.1f} us, sum of kernels {sum(e.time_range.end - e.time_range.start for e in evs):.1f} us, "
             f"{len(evs)} device activities, GPU fully idle {idle:.1f} us")
lines.append("per kernel: " + "; ".join(f"{k} x{c} {t:.0f}us" for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])))
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/timeline.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
