mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_o.log; tail -4 gpurun_out/pytest_o.log
python tools/gpu_profile_train.py bf16 > gpurun_out/train_profile_bf16.log 2>&1; head -14 gpurun_out/train_profile_bf16.log
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
torch.cuda.set_device(0)
print(bench.bench_finetune(torch.device('cuda:0'), 1, 20))
PY
