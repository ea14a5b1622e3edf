#!/bin/bash
# retry a gpurun call while the pod answers busy / transient (nothing is charged for those)
#   tools/gpurun_retry.sh <timeout_s> <command string> [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes\|rc=3"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
