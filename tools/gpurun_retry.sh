#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit code 3).  usage: tools/gpurun_retry.sh [gpurun args] -- 'command'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
