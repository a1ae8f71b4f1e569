#!/bin/bash
# tools/gpurun_retry.sh <log> <timeout> <command...> : retry a gpurun call while the pod answers "transient" (nothing is charged for those)
log=$1; to=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then exit 0; fi
  sleep 150
done
exit 3
