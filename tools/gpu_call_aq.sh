#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k gemm_ln 2>&1 | tail -25
echo "--- v2"; timeout 300 python tools/gpu_perf_gemm_ln.py 2>&1 | tee gpurun_out/gemm_ln2_ab.txt | tail -4
echo "--- v1"; BLM_GEMM_LN_V1=1 timeout 300 python tools/gpu_perf_gemm_ln.py 2>&1 | tee -a gpurun_out/gemm_ln2_ab.txt | tail -4
timeout 300 python tools/gpu_perf_gemm_ln.py 52833 2>&1 | tee -a gpurun_out/gemm_ln2_ab.txt | tail -4
