#!/bin/bash
# usage: gpurun_retry.sh <timeout> <logfile> <command...>   retries while the pod answers busy (rc 3), up to 10 times
T=$1; LOG=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
