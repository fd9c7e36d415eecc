#!/bin/bash
# GPU run 6: source-level ncu captures of the epilogue-bound GEMM shapes (K=1024) + A/B of the FFN1 bias-gradient placement.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export T=65536
ONLY="fwd ffn1 +bias+gelu" timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 10 -c 1 -o gpurun_out/r2_6_gemm_ffn1_gelu -f python tools/bench_gemm.py > gpurun_out/r2_6_ncu_a.log 2>&1; echo "ncu a rc=$?"
ONLY="dgrad ffn2 *gelu'(u) +colsum" timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 10 -c 1 -o gpurun_out/r2_6_gemm_dgelu -f python tools/bench_gemm.py > gpurun_out/r2_6_ncu_b.log 2>&1; echo "ncu b rc=$?"
ONLY="fwd qkv  [T,H]x[3H,H]" timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 10 -c 1 -o gpurun_out/r2_6_gemm_qkv -f python tools/bench_gemm.py > gpurun_out/r2_6_ncu_c.log 2>&1; echo "ncu c rc=$?"
ls -la gpurun_out/*.ncu-rep
for i in 1 2; do
  timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_6_bench_fused_$i.json 2> gpurun_out/r2_6_bench_fused_$i.err; echo "bench fused rc=$?"
  DLE_FFN1_BIAS_GRAD=separate timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_6_bench_sep_$i.json 2> gpurun_out/r2_6_bench_sep_$i.err; echo "bench separate rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_6_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
