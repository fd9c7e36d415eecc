#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <out_file> [--gpus N] <command>   -- retries while the pod answers "transient"/busy (nothing charged)
T=$1; OUT=$2; shift 2
EXTRA=""
if [ "$1" == "--gpus" ]; then EXTRA="--gpus $2"; shift 2; fi
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T $EXTRA -- "$@" > "$OUT" 2>&1
  rc=$?
  if grep -q "status=transient\|busy" "$OUT" && ! grep -q "status=ok" "$OUT"; then sleep 90; continue; fi
  if [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc attempt=$attempt" >> "$OUT"
