#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --tb=short -x 2>&1 | tail -n 8 > gpurun_out/pytest_backbone.log
tail -n 5 gpurun_out/pytest_backbone.log
bash tools/ab.sh "RT_CV_V2=1" "RT_CV_V2=0" 2>&1 | tee gpurun_out/ab_cv3.txt
