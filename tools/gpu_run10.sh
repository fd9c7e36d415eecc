#!/bin/bash
# GPU run 10: attention backward v5b (per-head lse/delta, smem loop bound, per-head bias atomics) tests + A/B vs v4c / v5 + trace.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=$PWD/deeplearningexamples_b200
timeout -k 10 300 python -m pytest tests/test_attention_gpu.py -m gpu -x -q > gpurun_out/r2_11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_11_pytest.log
for v in v4c v5b default v5b default; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  B=128 timeout -k 10 200 python tools/bench_attn.py 2>&1 | grep "p=0" | sed "s/^/attn $v: /"
done | tee gpurun_out/r2_11_attn_ab.log
unset DLE_LIB_PATH
B=128 timeout -k 10 200 python tools/attn_trace.py > gpurun_out/r2_11_attn_trace_v5c.log 2>&1; echo "trace rc=$?"; grep "pair period\|CTAs\|traced" gpurun_out/r2_11_attn_trace_v5b.log
cp gpurun_out/attn_trace.json gpurun_out/r2_11_attn_trace_v5c.json
exit 0
