#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k gemm_ln 2>&1 | tail -5
python tools/gpu_dbg_ln2.py 2>&1 | grep -v "^   " | head -4
echo "--- <2,4>"; timeout 300 python tools/gpu_perf_gemm_ln.py 52833 2>&1 | tail -3
echo "--- <3,2>"; BLM_GEMM_LN_32=1 timeout 300 python tools/gpu_perf_gemm_ln.py 52833 2>&1 | tail -3
echo "--- <2,4>"; timeout 300 python tools/gpu_perf_gemm_ln.py 65536 2>&1 | tail -3
