#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while gpurun answers "busy" (exit 3)
LOG=$1; TMO=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $attempt attempt(s)" >> $LOG; exit $rc; fi
  sleep 90
done
echo "gave up" >> $LOG; exit 3
