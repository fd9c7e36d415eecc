#!/bin/bash
# GPU run 8: attention backward with two MMA issuers (tests, A/B, trace) + GEMM epilogue without shared-memory staging (tests, A/B, step).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=$PWD/deeplearningexamples_b200
DLE_LIB_PATH=$P/libdle_b200_attn2.so timeout -k 10 300 python -m pytest tests/test_attention_gpu.py -m gpu -x -q > gpurun_out/r2_8_pytest_attn2.log 2>&1; echo "pytest attn2 rc=$?"; tail -2 gpurun_out/r2_8_pytest_attn2.log
for v in default attn2 default attn2; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  B=128 timeout -k 10 200 python tools/bench_attn.py 2>&1 | grep "mask,p=0.1" | sed "s/^/attn $v: /"
done | tee gpurun_out/r2_8_attn_ab.log
DLE_LIB_PATH=$P/libdle_b200_trace2.so B=128 timeout -k 10 200 python tools/attn_trace.py > gpurun_out/r2_8_attn_trace2.log 2>&1; echo "trace2 rc=$?"; head -12 gpurun_out/r2_8_attn_trace2.log
for v in dstore direct; do
  DLE_LIB_PATH=$P/libdle_b200_$v.so timeout -k 10 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r2_8_pytest_gemm_$v.log 2>&1; echo "pytest gemm $v rc=$?"; tail -2 gpurun_out/r2_8_pytest_gemm_$v.log
done
for v in default dstore direct default direct; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  T=65536 CASES=epi timeout -k 10 300 python tools/bench_gemm.py > gpurun_out/r2_8_gemm_$v.log 2>&1; echo "bench_gemm $v rc=$?"
  echo "== $v"; grep "case" gpurun_out/r2_8_gemm_$v.log | sed 's/nan/None/g' | python -c "
import sys, ast
for l in sys.stdin:
    d = ast.literal_eval(l.strip()); print('  %-44s %8.4f ms %7.1f TF' % (d['case'], d['ms'], d['tflops']))"
done 2>&1 | tee gpurun_out/r2_8_gemm_ab.log
for v in default direct attn2 default; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_8_bench_$v.json 2> gpurun_out/r2_8_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_8_bench_$v.json").read().strip().splitlines()[-1]); print("$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"])
except Exception as e: print("$v ERR", e)
PY
done
exit 0
