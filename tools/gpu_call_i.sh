mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv > gpurun_out/smi_i.txt
for i in 1 2; do
python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_i_stg_$i.log
BLM_NO_STG=1 python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_i_nostg_$i.log
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv >> gpurun_out/smi_i.txt
paste -d'\n' gpurun_out/perf_i_stg_1.log gpurun_out/perf_i_nostg_1.log | cut -c1-90
