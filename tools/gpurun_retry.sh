#!/bin/bash
# retry gpurun while the pod answers "busy" (exit 3); usage: gpurun_retry.sh TIMEOUT 'command'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
