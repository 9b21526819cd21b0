#!/bin/bash
# usage: retry_gpurun.sh LOG GPUS TIMEOUT SCRIPT — retries while the pod has no free slot (exit code 3), up to ~40 minutes
log=$1; gpus=$2; to=$3; script=$4
for i in $(seq 1 20); do
  if [ "$gpus" = "1" ]; then gpurun --timeout $to -- "bash $script" > $log 2>&1; else gpurun --gpus $gpus --timeout $to -- "bash $script" > $log 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
