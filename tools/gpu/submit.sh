#!/bin/bash
# usage: tools/gpu/submit.sh <gpus> <timeout_s> <script under tools/gpu> — retries while the pod answers "busy" (exit 3, nothing charged)
gpus=$1; to=$2; script=$3
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "mkdir -p gpurun_out; bash tools/gpu/$script > gpurun_out/${script%.sh}.log 2>&1; tail -5 gpurun_out/${script%.sh}.log"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
