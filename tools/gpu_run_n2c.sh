#!/bin/bash
# Same-box weak-scaling check: N=1 then N=2 (then N=1 again) on one 2-GPU box, default bench, CUDA graphs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2; do
  timeout -k 10 400 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_n2c_n1_$i.json 2> gpurun_out/r2_n2c_n1_$i.err; echo "n1 rc=$?"
  timeout -k 10 600 $TR --master-port 2951$i bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_n2c_n2_$i.json 2> gpurun_out/r2_n2c_n2_$i.err; echo "n2 rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_n2c_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"])
    except Exception as e: print(f, "ERR", e)
PY
exit 0
