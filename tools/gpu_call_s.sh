mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lstm.py -m gpu -q -x -k "cluster or lstm or Lstm or reparam or session or golden or injected or ragged" 2>&1 | tail -15 > gpurun_out/pytest_s.log; tail -6 gpurun_out/pytest_s.log
timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E -A1 "sampled|reparam|lstm" > gpurun_out/perf_s_cluster.log
BLM_SAMPLED_NO_CLUSTER=1 BLM_LSTM_NO_CLUSTER=1 timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E -A1 "sampled|reparam|lstm" > gpurun_out/perf_s_nocluster.log
cat gpurun_out/perf_s_cluster.log; echo ---; cat gpurun_out/perf_s_nocluster.log
