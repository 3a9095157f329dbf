mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transformer.py -m gpu -q -x -k "sampled or fused" 2>&1 | tail -4
BLM_SAMPLED_CLUSTER= python tools/gpu_perf_kernels.py 2>&1 | grep -E "sampled|reparam"
