#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py tests/test_gpu_reference_acceptance.py -m gpu -q --tb=short -x 2>&1 | tail -n 8
python tools/bench_ops.py > gpurun_out/ops_roofline.txt 2> gpurun_out/ops_roofline.err; cut -c1-130 gpurun_out/ops_roofline.txt
