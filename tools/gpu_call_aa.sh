mkdir -p gpurun_out
BLM_LSTM_CLUSTER= timeout 600 python - <<'PY' 2>&1 | tail -4
import os, subprocess, sys
env = dict(os.environ); env.pop("BLM_LSTM_CLUSTER", None)
print(subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_lstm.py", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"], env=dict(env, BLM_LSTM_CLUSTER="", BLM_LSTM_STAGGER=""), capture_output=True, text=True).stdout[-400:])
PY
for v in "" "BLM_LSTM_NO_SUB2=1"; do
echo "== $v"; env $v timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E -A1 "lstm_layer"
done
