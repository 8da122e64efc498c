#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).
#   tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
