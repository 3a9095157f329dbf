#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mn_major" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_train_loop.py -m gpu -x -q 2>&1 | tail -12
echo "--- MN-major"; timeout 300 python tools/gpu_profile_train.py bf16 2>&1 | grep -E "^step|transpose|split|bracketed" 
echo "--- transposes"; BLM_TRAIN_TRANSPOSE=1 timeout 300 python tools/gpu_profile_train.py bf16 2>&1 | grep -E "^step|transpose|split|bracketed"
