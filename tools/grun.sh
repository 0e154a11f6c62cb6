#!/bin/bash
# gpurun with retry while the pod has no free slot (rc 3 / "transient"): nothing is charged then.
# usage: tools/grun.sh <timeout-seconds> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes" || [ $rc -eq 3 ]; then
    sleep 90; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up after 40 retries"; exit 3
