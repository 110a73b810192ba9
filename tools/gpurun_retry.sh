#!/bin/bash
# gpurun with retries while the pod answers busy (rc 3 / transient): tools/gpurun_retry.sh <out> [--gpus N] -- <cmd>
OUT=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$OUT" 2>&1
  if grep -q "status=ok\|status=fail\|status=error\|status=timeout" "$OUT"; then exit 0; fi
  if ! grep -q "transient\|busy\|rc=3\|no box" "$OUT"; then exit 0; fi
  sleep 90
done
