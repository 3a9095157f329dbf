mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_n.log; tail -4 gpurun_out/pytest_n.log
python tools/gpu_profile_train.py bf16 > gpurun_out/train_profile_bf16.log 2>&1; cat gpurun_out/train_profile_bf16.log
python tools/gpu_profile_train.py bf16x3 > gpurun_out/train_profile_bf16x3.log 2>&1; head -12 gpurun_out/train_profile_bf16x3.log
