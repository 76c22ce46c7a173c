#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  usage: gpurun_retry.sh LOG TIMEOUT 'command'
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
