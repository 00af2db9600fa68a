#!/bin/bash
# usage: tools/gpurun_retry.sh <out.txt> <gpurun args...>   — retries while the pod answers busy/transient (exit code 3)
out=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$out"; then exit $rc; fi
  sleep 60
done
exit 3
