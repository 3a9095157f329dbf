mkdir -p gpurun_out
python tools/gpu_diag_fullsize.py > gpurun_out/diag_full.log 2>&1; echo "diag rc=$?"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "exit=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
