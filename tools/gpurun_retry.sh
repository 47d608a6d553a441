#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout_s> '<command>'   -- retries while the pod answers "busy" (status transient, nothing charged)
G=""
if [ "$1" == "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
  rc=$?
  if ! grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 30
done
exit 3
