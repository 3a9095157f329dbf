mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -m gpu -q -x 2>&1 | tail -4
for v in "" "BLM_LSTM_NO_STAGGER=1" "BLM_LSTM_NO_CLUSTER=1" "BLM_LSTM_NO_CLUSTER=1 BLM_LSTM_NO_STAGGER=1"; do
echo "== $v"; env $v timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E -A1 "lstm_layer T=20 B=2048|lstm_layer T=26"
done
