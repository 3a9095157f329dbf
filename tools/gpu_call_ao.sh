#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transformer.py -m gpu -x -q -k "sampled" 2>&1 | tail -15
timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -i "sampled\|reparam" | tee gpurun_out/sampled_once_ab.txt
