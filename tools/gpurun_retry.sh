#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged):  tools/gpurun_retry.sh LOG [gpurun args...]
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  echo "exit $rc (attempt $i)" >> "$LOG"
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
