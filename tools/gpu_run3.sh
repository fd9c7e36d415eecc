#!/bin/bash
# GPU run 3: everything after the MMA-issue restructure (uniform loops, elected issue, high warp ids) + wave-aware split-K + packed GELU.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -o timeout=400 -p no:cacheprovider --durations=8 > gpurun_out/r2_3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_3_pytest.log
tail -14 gpurun_out/r2_3_pytest.log
timeout -k 10 120 python tools/debug_squad.py 2>&1 | tail -6
for lib in "" nohint; do
  export DLE_LIB_PATH=$( [ -n "$lib" ] && echo $PWD/deeplearningexamples_b200/libdle_b200_$lib.so )
  echo "=== lib=${lib:-default}"
  B=128 timeout -k 10 300 python tools/bench_attn.py 2>&1 | tee gpurun_out/r2_3_attn_b128_${lib:-default}.log
  timeout -k 10 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_3_bench_${lib:-default}.json 2> gpurun_out/r2_3_bench_${lib:-default}.err; echo "bench rc=$?"
  grep -h "gemm {\|resident pass\|e2e pass" gpurun_out/r2_3_bench_${lib:-default}.err | head -16
done
unset DLE_LIB_PATH
B=32 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 8 -c 1 -o gpurun_out/r2_3_attn_bwd -f python tools/bench_attn.py > gpurun_out/r2_3_ncu_bwd.log 2>&1
B=32 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 12 -c 1 -o gpurun_out/r2_3_attn_fwd -f python tools/bench_attn.py > gpurun_out/r2_3_ncu_fwd.log 2>&1
for cfg in "32 --cuda-graphs" "64 " "64 --cuda-graphs"; do
  set -- $cfg; tag="b$1$( [ -n "$2" ] && echo _graphs )"
  timeout -k 10 600 python tools/bench_reference_gpu.py --arm reference --batch $1 --steps 6 $2 > gpurun_out/r2_3_ref_$tag.json 2> gpurun_out/r2_3_ref_$tag.err; echo "ref $tag rc=$?"; grep "host enqueue" gpurun_out/r2_3_ref_$tag.err; cut -c1-330 gpurun_out/r2_3_ref_$tag.json
done
timeout -k 10 600 python tools/bench_reference_gpu.py --arm ours --batch 32 --steps 6 --cuda-graphs > gpurun_out/r2_3_ours_via_ref_b32_graphs.json 2> gpurun_out/r2_3_ours_via_ref_b32_graphs.err; echo "ours-via-ref rc=$?"; grep "host enqueue" gpurun_out/r2_3_ours_via_ref_b32_graphs.err; cut -c1-330 gpurun_out/r2_3_ours_via_ref_b32_graphs.json
timeout -k 10 600 python tools/bench_infer.py > gpurun_out/r2_3_infer.log 2>&1; tail -6 gpurun_out/r2_3_infer.log | cut -c1-300
