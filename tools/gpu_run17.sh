#!/bin/bash
# GPU run 17: ncu --set full of the attention kernels in the benchmarked configuration (mask, dropout 0.1), B=32.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=32 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 58 -c 1 -o gpurun_out/r2_17_attn_bwd_drop -f python tools/bench_attn.py > gpurun_out/r2_17_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
B=32 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 64 -c 1 -o gpurun_out/r2_17_attn_fwd_drop -f python tools/bench_attn.py > gpurun_out/r2_17_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
ls -la gpurun_out/r2_17*.ncu-rep
exit 0
