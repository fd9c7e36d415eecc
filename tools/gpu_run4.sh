#!/bin/bash
# GPU run 4: fwd v3c / bwd nk-overlap / LN prefetch / no-hint default validation + microbenchmarks, driver-graph hang debug, remaining tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/debug_driver_graphs.sh 2>&1 | tail -30
timeout -k 10 900 python -m pytest tests -m gpu -q -o timeout=300 -p no:cacheprovider --durations=8 --deselect "tests/test_reference_driver_gpu.py::test_unmodified_reference_driver_trains_checkpoints_and_resumes_over_the_b200_mirror[True]" > gpurun_out/r2_4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_4_pytest.log
tail -14 gpurun_out/r2_4_pytest.log | cut -c1-200
B=128 timeout -k 10 300 python tools/bench_attn.py 2>&1 | tee gpurun_out/r2_4_attn_b128.log
DLE_LIB_PATH=$PWD/deeplearningexamples_b200/libdle_b200_hint.so B=128 timeout -k 10 300 python tools/bench_attn.py 2>&1 | tee gpurun_out/r2_4_attn_b128_hint.log
T=65536 timeout -k 10 300 python tools/bench_ln.py 2>&1 | tee gpurun_out/r2_4_ln.log
timeout -k 10 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_4_bench.json 2> gpurun_out/r2_4_bench.err; echo "bench rc=$?"
grep -h "gemm {\|resident pass\|e2e pass" gpurun_out/r2_4_bench.err | head -14
timeout -k 10 600 python bench.py --workload squad --steps 8 --warmup 3 > gpurun_out/r2_4_bench_squad.json 2> gpurun_out/r2_4_bench_squad.err; echo "squad rc=$?"; grep -h "resident pass\|e2e pass\|capture" gpurun_out/r2_4_bench_squad.err; cut -c1-300 gpurun_out/r2_4_bench_squad.json
timeout -k 10 600 python bench.py --seq 128 --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_4_bench_s128.json 2> gpurun_out/r2_4_bench_s128.err; echo "s128 rc=$?"; grep -h "resident pass\|e2e pass" gpurun_out/r2_4_bench_s128.err
timeout -k 10 600 python tools/bench_infer.py > gpurun_out/r2_4_infer.log 2>&1; grep "^{'batch" gpurun_out/r2_4_infer.log | cut -c1-220
timeout -k 10 300 python tools/bench_lamb.py > gpurun_out/r2_4_lamb.log 2>&1; tail -4 gpurun_out/r2_4_lamb.log | cut -c1-250
