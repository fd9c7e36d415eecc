#!/bin/bash
# GPU run 16: full GPU suite + smoke + a bench line on the final tree (GEMM waits hinted).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_16_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_16_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_16_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_16_smoke.log
timeout -k 10 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_16_bench.json 2> gpurun_out/r2_16_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_16_bench.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"])
PY
exit 0
