mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
echo "== perf TMA store (default)"; timeout 300 python tools/gpu_perf_kernels.py 2>&1 | sed -n 2,9p
echo "== perf BLM_TMA_STORE=0"; BLM_TMA_STORE=0 timeout 300 python tools/gpu_perf_kernels.py 2>&1 | sed -n 2,9p
