#!/bin/bash
# 2-GPU run: NCCL equivalence tests of the product path, DDP bench with the whole step captured as a CUDA graph vs eager, reference arm at N=2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 10 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -o timeout=400 -p no:cacheprovider > gpurun_out/r2_n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 900 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_n2_bench_graphs.json 2> gpurun_out/r2_n2_bench_graphs.err; echo "graphs rc=$?"
grep -h "captured\|capture failed\|resident pass\|e2e pass\|nccl:" gpurun_out/r2_n2_bench_graphs.err | tail -20; cut -c1-400 gpurun_out/r2_n2_bench_graphs.json
timeout -k 10 900 $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-cuda-graphs > gpurun_out/r2_n2_bench_eager.json 2> gpurun_out/r2_n2_bench_eager.err; echo "eager rc=$?"
grep -h "resident pass\|e2e pass" gpurun_out/r2_n2_bench_eager.err; cut -c1-300 gpurun_out/r2_n2_bench_eager.json
timeout -k 10 900 $TR --master-port 29513 tools/bench_reference_gpu.py --arm reference --batch 32 --steps 6 > gpurun_out/r2_n2_ref_b32.json 2> gpurun_out/r2_n2_ref_b32.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_n2_ref_b32.json
timeout -k 10 900 $TR --master-port 29514 tools/bench_reference_gpu.py --arm reference --batch 64 --steps 6 --cuda-graphs > gpurun_out/r2_n2_ref_b64g.json 2> gpurun_out/r2_n2_ref_b64g.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_n2_ref_b64g.json
