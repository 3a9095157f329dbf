mkdir -p gpurun_out
python tools/gpu_diag_kernels.py > gpurun_out/diag_ew8.log 2>&1; echo "diag ew8 rc=$?"
BLM_EPI_WARPS=16 python tools/gpu_diag_kernels.py > gpurun_out/diag_ew16.log 2>&1; echo "diag ew16 rc=$?"
python tools/gpu_perf_kernels.py > gpurun_out/perf_ew8.log 2>&1; echo "perf ew8 rc=$?"
BLM_EPI_WARPS=16 python tools/gpu_perf_kernels.py > gpurun_out/perf_ew16.log 2>&1; echo "perf ew16 rc=$?"
grep -c OK gpurun_out/diag_ew8.log gpurun_out/diag_ew16.log; grep -h "FAIL" gpurun_out/diag_ew8.log gpurun_out/diag_ew16.log | head
