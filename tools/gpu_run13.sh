#!/bin/bash
# GPU run 13: attention A/B of the register-cap change, then the round's evidence on the final tree: attention tests, default bench (CPU baseline +
# reference GPU leg), ncu launch list, ncu --set full of the attention kernels and of 12 forward GEMM launches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=$PWD/deeplearningexamples_b200
timeout -k 10 300 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2_13_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_13_pytest.log
for v in default; do
  if [ $v = default ]; then unset DLE_LIB_PATH; else export DLE_LIB_PATH=$P/libdle_b200_$v.so; fi
  B=128 timeout -k 10 200 python tools/bench_attn.py 2>&1 | grep "p=0" | sed "s/^/attn $v: /"
done | tee gpurun_out/r2_13_attn_ab.log
unset DLE_LIB_PATH
B=128 timeout -k 10 200 python tools/bench_attn.py > gpurun_out/r2_13_attn_b128.log 2>&1; cp gpurun_out/bench_attn.json gpurun_out/r2_13_attn_b128.json
timeout -k 10 200 python tools/bench_attn.py > gpurun_out/r2_13_attn_b32.log 2>&1
timeout -k 10 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_13_bench.json 2> gpurun_out/r2_13_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_13_bench.json
B=32 timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 8 -c 1 -o gpurun_out/r2_13_attn_bwd -f python tools/bench_attn.py > gpurun_out/r2_13_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
B=32 timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 12 -c 1 -o gpurun_out/r2_13_attn_fwd -f python tools/bench_attn.py > gpurun_out/r2_13_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1400 --csv --log-file gpurun_out/r2_13_launches.csv python bench.py --steps 2 --warmup 3 --no-cuda-graphs --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_13_ncu_launch_bench.json 2> gpurun_out/r2_13_ncu_launch_bench.err; echo "launch list rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_bf16_tcgen05 -c 12 --csv --page raw --log-file gpurun_out/r2_13_prof_gemm_raw.csv python bench.py --steps 1 --warmup 3 --no-cuda-graphs --no-cpu-baseline --no-reference-gpu > /dev/null 2> gpurun_out/r2_13_ncu_gemm.err; echo "gemm capture rc=$?"
exit 0
