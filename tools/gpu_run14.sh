#!/bin/bash
# GPU run 14: final-tree numbers of the other workloads (phase 1, SQuAD step, encoder-only inference) + the full GPU suite once more.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python bench.py --workload squad --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_14_bench_squad.json 2> gpurun_out/r2_14_bench_squad.err; echo "squad rc=$?"; grep -h "resident pass\|e2e pass\|capture" gpurun_out/r2_14_bench_squad.err
timeout -k 10 600 python bench.py --seq 128 --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_14_bench_s128.json 2> gpurun_out/r2_14_bench_s128.err; echo "s128 rc=$?"; grep -h "resident pass\|e2e pass" gpurun_out/r2_14_bench_s128.err
timeout -k 10 600 python tools/bench_infer.py > gpurun_out/r2_14_infer.log 2>&1; grep "^{'batch" gpurun_out/r2_14_infer.log | cut -c1-220
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_14_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_14_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_14_smoke.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_14_bench_squad.json", "gpurun_out/r2_14_bench_s128.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["model_flops_utilisation"])
    except Exception as e: print(f, "ERR", e)
PY
exit 0
