#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] <command>   - retries while the pod answers "busy" (exit code 3 / transient)
T=$1; shift
OPTS=""
if [ "$1" == "--gpus" ]; then OPTS="--gpus $2"; shift; shift; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout "$T" $OPTS -- "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$rc" != "3" ] && [ "$st" != "transient" ]; then exit $rc; fi
  echo "[retry] attempt $i busy; sleeping 120 s"; sleep 120
done
exit 3
