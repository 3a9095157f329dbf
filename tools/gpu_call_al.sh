#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "cta_pair" 2>&1 | tail -15
echo "--- BLM_GEMM2=0"; timeout 300 python tools/gpu_perf_kernels.py 2>&1 | head -12 | tee gpurun_out/gemm2_ab.txt
echo "--- BLM_GEMM2=1 (TMA store)"; BLM_GEMM2=1 timeout 300 python tools/gpu_perf_kernels.py 2>&1 | head -12 | tee -a gpurun_out/gemm2_ab.txt
echo "--- BLM_GEMM2=1 BLM_GEMM2_TMA_STORE=0"; BLM_GEMM2=1 BLM_GEMM2_TMA_STORE=0 timeout 300 python tools/gpu_perf_kernels.py 2>&1 | head -12 | tee -a gpurun_out/gemm2_ab.txt
