#!/bin/bash
# 2-GPU validation of the final tree: NCCL equivalence tests of the product path + DDP bench (whole step captured as one CUDA graph per rank).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 10 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -o timeout=400 -p no:cacheprovider > gpurun_out/r2_n2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_n2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 600 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_n2b_bench_graphs.json 2> gpurun_out/r2_n2b_bench_graphs.err; echo "graphs rc=$?"
grep -h "captured\|capture failed\|resident pass\|e2e pass\|nccl:" gpurun_out/r2_n2b_bench_graphs.err | tail -12; cut -c1-500 gpurun_out/r2_n2b_bench_graphs.json
exit 0
