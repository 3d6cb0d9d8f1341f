#!/bin/bash
# measurement aid: run a gpurun call, retrying while the pod answers "busy" (exit 3 / status=transient, nothing charged)
#   tools/gpurun_retry.sh <timeout-seconds> '<command>'
T=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  if [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gpurun: still busy after 12 attempts"; exit 3
