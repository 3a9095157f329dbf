mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lstm.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_m.log; tail -6 gpurun_out/pytest_m.log
python tools/gpu_perf_kernels.py 2>&1 | grep -A1 lstm_layer > gpurun_out/perf_m_u16.log
BLM_LSTM_U8=1 python tools/gpu_perf_kernels.py 2>&1 | grep -A1 lstm_layer > gpurun_out/perf_m_u8.log
cat gpurun_out/perf_m_u16.log; echo ---; cat gpurun_out/perf_m_u8.log
