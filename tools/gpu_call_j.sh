mkdir -p gpurun_out
python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_j_auto.log
BLM_STG=0 python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_j_nostg.log
BLM_STG=1 python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_j_stg.log
paste -d'\n' gpurun_out/perf_j_auto.log gpurun_out/perf_j_nostg.log gpurun_out/perf_j_stg.log | cut -c1-90
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_j.log; tail -5 gpurun_out/pytest_j.log
python tools/profile_step.py > gpurun_out/plain_r01j.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 27 -c 1 -f -o gpurun_out/prof_r01j_ffn1 python tools/profile_step.py > gpurun_out/ncu_r01j.log 2>&1
tail -2 gpurun_out/ncu_r01j.log
