#!/bin/bash
# GPU run 2: attention v4 validation + microbench + ncu, reference-driver tests (fixed), new loss/squad tests, reference GPU arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_large_width_gpu.py -m gpu -q -x -o timeout=300 -p no:cacheprovider > gpurun_out/r2_2_pytest_attn.log 2>&1; echo "attn pytest rc=$?"
tail -4 gpurun_out/r2_2_pytest_attn.log
B=32 timeout -k 10 300 python tools/bench_attn.py > gpurun_out/r2_2_attn_b32.log 2>&1; cat gpurun_out/r2_2_attn_b32.log
B=128 timeout -k 10 300 python tools/bench_attn.py > gpurun_out/r2_2_attn_b128.log 2>&1; cat gpurun_out/r2_2_attn_b128.log
B=32 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 8 -c 1 -o gpurun_out/r2_2_attn_bwd -f python tools/bench_attn.py > gpurun_out/r2_2_ncu_bwd.log 2>&1
B=32 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 12 -c 1 -o gpurun_out/r2_2_attn_fwd -f python tools/bench_attn.py > gpurun_out/r2_2_ncu_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout -k 10 900 python -m pytest tests/test_reference_driver_gpu.py tests/test_loss_gpu.py tests/test_squad_gpu.py -m gpu -q -o timeout=400 -p no:cacheprovider --durations=12 > gpurun_out/r2_2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_2_pytest.log
tail -25 gpurun_out/r2_2_pytest.log
for cfg in "32 " "32 --cuda-graphs" "64 --cuda-graphs"; do
  set -- $cfg; tag="b$1$( [ -n "$2" ] && echo _graphs )"
  timeout -k 10 600 python tools/bench_reference_gpu.py --arm reference --batch $1 --steps 6 $2 > gpurun_out/r2_2_ref_$tag.json 2> gpurun_out/r2_2_ref_$tag.err; echo "ref $tag rc=$?"; tail -2 gpurun_out/r2_2_ref_$tag.err; cat gpurun_out/r2_2_ref_$tag.json
done
timeout -k 10 600 python tools/bench_reference_gpu.py --arm ours --batch 32 --steps 6 --cuda-graphs > gpurun_out/r2_2_ours_via_ref_b32_graphs.json 2> gpurun_out/r2_2_ours_via_ref_b32_graphs.err; echo "ours-via-ref rc=$?"; tail -3 gpurun_out/r2_2_ours_via_ref_b32_graphs.err; cat gpurun_out/r2_2_ours_via_ref_b32_graphs.json
timeout -k 10 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err; echo "bench rc=$?"; grep -h "gemm {\|host enqueue\|resident pass\|e2e pass" gpurun_out/r2_2_bench.err
