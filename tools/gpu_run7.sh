#!/bin/bash
# GPU run 7: GEMM epilogue-warp A/B (8 warps x 4 stages | 8 x 3 | 16 x 3) + attention-backward hand-off trace.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=$PWD/deeplearningexamples_b200
DLE_LIB_PATH=$P/libdle_b200_epi16.so timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r2_7_pytest_gemm_epi16.log 2>&1; echo "pytest gemm epi16 rc=$?"; tail -3 gpurun_out/r2_7_pytest_gemm_epi16.log
timeout -k 10 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r2_7_pytest_gemm_default.log 2>&1; echo "pytest gemm default rc=$?"; tail -2 gpurun_out/r2_7_pytest_gemm_default.log
for v in default st3 epi16 default epi16; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  T=65536 CASES=epi timeout -k 10 300 python tools/bench_gemm.py > gpurun_out/r2_7_gemm_$v.log 2>&1; echo "bench_gemm $v rc=$?"
  cp gpurun_out/r2_7_gemm_$v.log gpurun_out/r2_7_gemm_${v}_$(date +%s).log
done
unset DLE_LIB_PATH
B=128 timeout -k 10 300 python tools/attn_trace.py > gpurun_out/r2_7_attn_trace.log 2>&1; echo "attn trace rc=$?"; head -30 gpurun_out/r2_7_attn_trace.log
for v in default epi16; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_7_bench_$v.json 2> gpurun_out/r2_7_bench_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_7_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"])
    except Exception as e: print(f, "ERR", e)
PY
for v in default st3 epi16; do echo "== $v"; grep "case" gpurun_out/r2_7_gemm_$v.log | python -c "
import sys, ast
for l in sys.stdin:
    d = ast.literal_eval(l.strip()); print('  %-44s %8.4f ms %7.1f TF' % (d['case'], d['ms'], d['tflops']))"; done
