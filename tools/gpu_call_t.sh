mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E -A1 "lstm" > gpurun_out/perf_t.log
BLM_LSTM_NO_CLUSTER=1 timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E -A1 "lstm" > gpurun_out/perf_t_nocluster.log
cat gpurun_out/perf_t.log; echo ---; cat gpurun_out/perf_t_nocluster.log
