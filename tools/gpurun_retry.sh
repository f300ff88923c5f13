#!/bin/bash
# usage: gpurun_retry.sh LOG [gpurun args...] — retries while the pod answers busy/transient (nothing is charged for those)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 45
done
exit 3
