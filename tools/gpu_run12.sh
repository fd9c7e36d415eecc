#!/bin/bash
# GPU run 12: full GPU suite on the new kernels; attention A/B v5c vs v5d (prefetched row statistics, drain deferred into the next pair,
# coalesced forward O store and delta kernel); trace; step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=$PWD/deeplearningexamples_b200
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_12_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_12_pytest.log
for v in v5c default v5c default; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  B=128 timeout -k 10 200 python tools/bench_attn.py 2>&1 | grep "p=0" | sed "s/^/attn $v: /"
done | tee gpurun_out/r2_12_attn_ab.log
unset DLE_LIB_PATH
B=128 timeout -k 10 200 python tools/attn_trace.py > gpurun_out/r2_12_attn_trace_v5d.log 2>&1; echo "trace rc=$?"; grep "pair period\|CTAs\|traced" gpurun_out/r2_12_attn_trace_v5d.log
cp gpurun_out/attn_trace.json gpurun_out/r2_12_attn_trace_v5d.json
for v in default v5c default; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_12_bench_$v.json 2> gpurun_out/r2_12_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_12_bench_$v.json").read().strip().splitlines()[-1]); print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"])
except Exception as e: print("$v ERR", e)
PY
done
exit 0
