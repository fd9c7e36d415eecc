#!/bin/bash
# GPU run 15: A/B of suspend-time hints on the GEMM kernel's mbarrier waits only (attention keeps the un-hinted form).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=$PWD/deeplearningexamples_b200
for v in default gemmhint default gemmhint; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  timeout -k 10 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_15_bench_$v.json 2> gpurun_out/r2_15_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_15_bench_$v.json").read().strip().splitlines()[-1]); print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], d["roofline"]["achieved"])
except Exception as e: print("$v ERR", e)
PY
done
exit 0
