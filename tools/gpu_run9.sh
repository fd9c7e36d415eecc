#!/bin/bash
# GPU run 9: attention backward v5 (K/V double buffer, dO ring) tests + A/B vs v4c + trace; GEMM tests after the EW template change; step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=$PWD/deeplearningexamples_b200
timeout -k 10 300 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r2_9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_9_pytest.log
for v in v4c default v4c default; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  B=128 timeout -k 10 200 python tools/bench_attn.py 2>&1 | grep "p=0" | sed "s/^/attn $v: /"
done | tee gpurun_out/r2_9_attn_ab.log
unset DLE_LIB_PATH
B=128 timeout -k 10 200 python tools/attn_trace.py > gpurun_out/r2_9_attn_trace_v5.log 2>&1; echo "trace rc=$?"; grep "pair period\|CTAs\|traced" gpurun_out/r2_9_attn_trace_v5.log
T=65536 CASES=epi timeout -k 10 300 python tools/bench_gemm.py 2>&1 | grep "case" | sed 's/nan/None/g' | python -c "
import sys, ast
for l in sys.stdin:
    d = ast.literal_eval(l.strip()); print('  %-44s %8.4f ms %7.1f TF' % (d['case'], d['ms'], d['tflops']))" | tee gpurun_out/r2_9_gemm.log
for v in default v4c default; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_9_bench_$v.json 2> gpurun_out/r2_9_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_9_bench_$v.json").read().strip().splitlines()[-1]); print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"])
except Exception as e: print("$v ERR", e)
PY
done
exit 0
