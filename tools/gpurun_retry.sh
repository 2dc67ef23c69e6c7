#!/bin/bash
# usage: tools/gpurun_retry.sh <logname> <timeout_s> <command...>   -- retries while the pod answers "busy" (exit 3)
LOG=gpurun_out/$1; shift; TMO=$1; shift
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then echo "rc=$rc" >> $LOG; exit $rc; fi
  sleep 60
done
echo "gave up" >> $LOG
