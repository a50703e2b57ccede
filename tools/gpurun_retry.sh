#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers "busy"
T=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    sleep 120
    continue
  fi
  echo "$out" | tail -60
  exit 0
done
echo "gave up after 40 busy answers"
