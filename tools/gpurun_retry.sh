#!/bin/bash
# usage: tools/gpurun_retry.sh TAG [--gpus N] SCRIPT -- retries while the pod answers "busy" (exit 3)
tag=$1; shift
for i in $(seq 1 30); do
  gpurun --timeout 2400 "${@:1:$#-1}" -- bash "${@: -1}" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "nothing was charged" gpurun_out/${tag}_call.log; then break; fi
  sleep 150
done
echo "done rc=$rc try=$i" >> gpurun_out/${tag}_call.log
