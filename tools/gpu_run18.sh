#!/bin/bash
# GPU run 18: the GEMM tests incl. the new ragged / wide cases of the 16-warp gelu'(u) instance.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_gemm_gpu.py -m gpu -q > gpurun_out/r2_18_pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_18_pytest_gemm.log
exit 0
