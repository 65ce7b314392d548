#!/bin/bash
# usage: [GPUS=N] retry.sh <timeout> <command>  -- gpurun, retried while the pod answers "busy" (exit code 3)
T=$1; shift
G=""
if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
