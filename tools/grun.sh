#!/bin/bash
# usage: tools/grun.sh <tag> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy; log in gpurun_out/<tag>_call.log)
tag=$1; to=$2; shift 2
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> gpurun_out/${tag}_call.log; exit $rc; fi
  sleep 90
done
