#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_train_loop.py tests/test_gpu_kernels.py -m gpu -x -q -k "train or evaluate or schedule or fast_gelu" 2>&1 | tail -30
timeout 300 python tools/gpu_perf_kernels.py 2>&1 | head -8
