#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <out_file> <command...>   -- retries while the pod answers "transient"/busy (nothing charged)
T=$1; OUT=$2; shift 2
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > "$OUT" 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|busy" "$OUT" && ! grep -q "status=ok" "$OUT"; then sleep 90; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc attempt=$attempt" >> "$OUT"
