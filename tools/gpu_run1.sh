#!/bin/bash
# GPU run 1 of round 2: full -m gpu suite, attention / LN microbenchmarks, bench.py (graphs + reference GPU leg), eager bench, drop-in path.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_1_smi.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q -o timeout=300 -p no:cacheprovider > gpurun_out/r2_1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_1_pytest.log
tail -5 gpurun_out/r2_1_pytest.log
B=32 timeout -k 10 300 python tools/bench_attn.py > gpurun_out/r2_1_attn_b32.log 2>&1; cp gpurun_out/bench_attn.json gpurun_out/r2_1_attn_b32.json 2>/dev/null
B=128 timeout -k 10 300 python tools/bench_attn.py > gpurun_out/r2_1_attn_b128.log 2>&1; cp gpurun_out/bench_attn.json gpurun_out/r2_1_attn_b128.json 2>/dev/null
cat gpurun_out/r2_1_attn_b128.log
T=65536 timeout -k 10 300 python tools/bench_ln.py > gpurun_out/r2_1_ln.log 2>&1
timeout -k 10 1200 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_1_bench.json
timeout -k 10 600 python bench.py --steps 6 --warmup 3 --no-cuda-graphs --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_1_bench_eager.json 2> gpurun_out/r2_1_bench_eager.err; echo "eager rc=$?"
timeout -k 10 600 python tools/bench_reference_gpu.py --arm ours --batch 32 --steps 6 --cuda-graphs > gpurun_out/r2_1_ours_via_ref_b32_graphs.json 2> gpurun_out/r2_1_ours_via_ref_b32_graphs.err; echo "ours-via-ref rc=$?"
grep -h "host enqueue\|resident pass\|e2e pass" gpurun_out/r2_1_bench.err gpurun_out/r2_1_bench_eager.err gpurun_out/r2_1_ours_via_ref_b32_graphs.err
