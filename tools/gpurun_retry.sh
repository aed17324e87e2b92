#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "busy" (exit 3, uncharged)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
