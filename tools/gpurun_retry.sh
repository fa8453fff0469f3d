#!/bin/bash
# usage: tools/gpurun_retry.sh <out.txt> <gpurun args...>   — retries while the pod answers busy (exit 3)
out=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$out" 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "retry in a few minutes" "$out"; then exit $rc; fi
  sleep 90
done
