mkdir -p gpurun_out
BLM_GEMM2=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pair_kernel or vocab_nll or fast_gelu" 2>&1 | tail -12
echo "== perf default"; timeout 300 python tools/gpu_perf_kernels.py 2>&1 | sed -n 2,13p
echo "== perf BLM_GEMM2=1"; BLM_GEMM2=1 timeout 300 python tools/gpu_perf_kernels.py 2>&1 | sed -n 2,13p
