#!/bin/bash
# usage: tools/gpu/retry.sh <log> <timeout> [--gpus N] -- <command>   (retries while gpurun answers "busy", rc 3)
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
