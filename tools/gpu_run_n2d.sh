#!/bin/bash
# N=2 A/B of NCCL's CTA budget (32 channels by default: 32 spinning CTAs per allreduce next to a 148-CTA persistent GEMM).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
i=0
for c in default 8 4 default 8; do
  i=$((i+1))
  if [ $c = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  timeout -k 10 600 $TR --master-port 2952$i bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_n2d_${i}_$c.json 2> gpurun_out/r2_n2d_${i}_$c.err; echo "ctas=$c rc=$?"
  grep -h "nccl: AllReduce: 1381" gpurun_out/r2_n2d_${i}_$c.err | head -1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_n2d_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"])
    except Exception as e: print(f, "ERR", e)
PY
exit 0
