mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_u.log; tail -3 gpurun_out/pytest_u.log
python tools/gpu_profile_train.py bf16 2>&1 | head -8
python tools/profile_lstm.py > gpurun_out/plain_lstm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lstm_layer -s 1 -c 1 -f -o gpurun_out/prof_r01u_lstm python tools/profile_lstm.py > gpurun_out/ncu_r01u.log 2>&1
tail -2 gpurun_out/ncu_r01u.log
