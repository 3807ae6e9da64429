#!/bin/bash
# usage: gpu_retry.sh <outfile> <timeout> <gpus> <command...>
out=$1; shift; to=$1; shift; gpus=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $out 2>&1; else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > $out 2>&1; fi
  rc=$?
  if grep -q "status=transient\|status=busy" $out || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
