#!/bin/bash
# GPU run 2: reference-driver tests (fixed), new loss/squad tests, ncu captures of the attention kernels, reference GPU arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_reference_driver_gpu.py tests/test_loss_gpu.py tests/test_squad_gpu.py -m gpu -q -o timeout=400 -p no:cacheprovider --durations=12 > gpurun_out/r2_2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_2_pytest.log
tail -25 gpurun_out/r2_2_pytest.log
B=32 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 8 -c 1 -o gpurun_out/r2_2_attn_bwd -f python tools/bench_attn.py > gpurun_out/r2_2_ncu_bwd.log 2>&1
B=32 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 12 -c 1 -o gpurun_out/r2_2_attn_fwd -f python tools/bench_attn.py > gpurun_out/r2_2_ncu_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout -k 10 600 python tools/bench_reference_gpu.py --arm reference --batch 32 --steps 6 > gpurun_out/r2_2_ref_b32_eager.json 2> gpurun_out/r2_2_ref_b32_eager.err; echo "ref eager rc=$?"; tail -3 gpurun_out/r2_2_ref_b32_eager.err; cat gpurun_out/r2_2_ref_b32_eager.json
timeout -k 10 600 python tools/bench_reference_gpu.py --arm reference --batch 32 --steps 6 --cuda-graphs > gpurun_out/r2_2_ref_b32_graphs.json 2> gpurun_out/r2_2_ref_b32_graphs.err; echo "ref graphs rc=$?"; tail -3 gpurun_out/r2_2_ref_b32_graphs.err; cat gpurun_out/r2_2_ref_b32_graphs.json
timeout -k 10 600 python tools/bench_reference_gpu.py --arm reference --batch 64 --steps 6 --cuda-graphs > gpurun_out/r2_2_ref_b64_graphs.json 2> gpurun_out/r2_2_ref_b64_graphs.err; echo "ref b64 rc=$?"; tail -3 gpurun_out/r2_2_ref_b64_graphs.err; cat gpurun_out/r2_2_ref_b64_graphs.json
timeout -k 10 600 python tools/bench_reference_gpu.py --arm ours --batch 32 --steps 6 --cuda-graphs > gpurun_out/r2_2_ours_via_ref_b32_graphs.json 2> gpurun_out/r2_2_ours_via_ref_b32_graphs.err; echo "ours-via-ref rc=$?"; tail -3 gpurun_out/r2_2_ours_via_ref_b32_graphs.err; cat gpurun_out/r2_2_ours_via_ref_b32_graphs.json
