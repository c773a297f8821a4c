#!/bin/bash
# usage: profiles/gpurun_retry.sh <timeout-seconds> <logfile> <command...>   -- retries while the pod answers "busy" (rc 3)
t=$1; log=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
