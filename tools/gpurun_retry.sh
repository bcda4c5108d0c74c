#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-s> '<command>'   -- retries while the pod answers "busy" (rc 3), up to ~40 min
T=$1; shift
for try in $(seq 1 20); do
    /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 60
done
exit 3
