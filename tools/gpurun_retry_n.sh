#!/bin/bash
# usage: tools/gpurun_retry_n.sh <n_gpus> <timeout_s> <command...>   — gpurun --gpus N, retried while the pod answers "busy" (exit 3)
n=$1; t=$2; shift; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $t -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
