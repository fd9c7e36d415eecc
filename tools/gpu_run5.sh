#!/bin/bash
# GPU run 5: full -m gpu suite on the final tree, LN microbench (cp.async ring), the default bench (CPU baseline + reference GPU leg), ncu launch list + GEMM traffic capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/r2_5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_5_pytest.log
tail -12 gpurun_out/r2_5_pytest.log | cut -c1-200
T=65536 timeout -k 10 300 python tools/bench_ln.py 2>&1 | tee gpurun_out/r2_5_ln.log
timeout -k 10 1500 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_5_bench.json 2> gpurun_out/r2_5_bench.err; echo "bench rc=$?"
grep -h "resident pass\|e2e pass\|reference gpu leg\|cpu baseline" gpurun_out/r2_5_bench.err | cut -c1-250
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1400 --csv --log-file gpurun_out/r2_5_launches.csv python bench.py --steps 2 --warmup 3 --no-cuda-graphs --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_5_ncu_launch_bench.json 2> gpurun_out/r2_5_ncu_launch_bench.err; echo "launch list rc=$?"; wc -l gpurun_out/r2_5_launches.csv
timeout -k 10 900 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_bf16_tcgen05 -c 12 --csv --page raw --log-file gpurun_out/r2_5_prof_gemm_raw.csv python bench.py --steps 1 --warmup 3 --no-cuda-graphs --no-cpu-baseline --no-reference-gpu > /dev/null 2> gpurun_out/r2_5_ncu_gemm.err; echo "gemm capture rc=$?"; wc -l gpurun_out/r2_5_prof_gemm_raw.csv
