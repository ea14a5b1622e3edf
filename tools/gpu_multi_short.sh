#!/bin/bash
# N-GPU check (N = $1) of the bench exactly as the driver launches it (its line carries the `train` / `train_cfg4` records too)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_host.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cut -c1-600 gpurun_out/bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
for k in ("train", "train_cfg4"):
    t = d.get(k, {})
    print(k, {q: t.get(q) for q in ("value", "ms_per_step", "n_gpus", "peak_mem_gb", "unavailable")}, t.get("collectives"))
PY
