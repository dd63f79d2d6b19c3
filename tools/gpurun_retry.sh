#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing is charged for those); usage: gpurun_retry.sh <timeout> <command...>
to=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$to" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|rc=3\|no box"; then echo "[retry $attempt] pod busy"; sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
