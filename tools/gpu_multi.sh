#!/bin/bash
# N-GPU check (N = $1): the bench exactly as the driver launches it, the reference arm, and the sharded training step
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_host.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --train --steps 3 --warmup 2 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-700; tail -n 3 gpurun_out/bench_n$N.err
cat gpurun_out/bench_train_n$N.json | cut -c1-500; tail -n 3 gpurun_out/bench_train_n$N.err
