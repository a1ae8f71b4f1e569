#!/bin/bash
# tools/gpurun_retry.sh <log> <timeout> <command...> : retry a gpurun call while the pod answers "transient" (nothing is charged for those)
# GPURUN_GPUS=N in the environment asks for an N-GPU box.
log=$1; to=$2; shift 2
extra=""
if [ -n "$GPURUN_GPUS" ]; then extra="--gpus $GPURUN_GPUS"; fi
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $extra --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then exit 0; fi
  sleep 150
done
exit 3
