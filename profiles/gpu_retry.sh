#!/bin/bash
# gpu_retry.sh [gpurun options] -- 'command' : retry while the pod answers "busy / transient" (nothing charged).
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then
    echo "[gpu_retry] attempt $attempt: busy, sleeping"; sleep 90; continue
  fi
  exit 0
done
exit 3
