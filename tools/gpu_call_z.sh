mkdir -p gpurun_out
python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_z_auto.log
BLM_STG=1 python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_z_stg.log
paste -d'\n' gpurun_out/perf_z_auto.log gpurun_out/perf_z_stg.log | cut -c1-95
