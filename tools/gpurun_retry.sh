#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...>   — retries while the pod answers "transient" (nothing charged)
t=$1; shift
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry $attempt] transient, sleeping 90 s"; sleep 90; continue
  fi
  echo "$out"; exit 0
done
echo "gave up after 30 transient answers"; exit 3
