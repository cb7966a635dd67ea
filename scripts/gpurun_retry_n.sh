#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <timeout> <logfile> <command...>   like gpurun_retry.sh with --gpus N
N=$1; T=$2; LOG=$3; shift 3
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
