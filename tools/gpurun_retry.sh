#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...] -- 'command'    (retries while the pod answers "transient")
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry $i] pod busy, sleeping 150 s"; sleep 150
  else
    break
  fi
done
