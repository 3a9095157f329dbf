mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "exit=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python tools/gpu_perf_kernels.py > gpurun_out/perf_g.log 2>&1; echo "perf rc=$?"
BLM_GEMM_NO_ARES=1 python tools/gpu_perf_kernels.py 2>&1 | head -8 > gpurun_out/perf_g_noares.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
python tools/profile_step.py > gpurun_out/plain_r01g.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 49 -c 1 -f -o gpurun_out/prof_r01g_nll python tools/profile_step.py > gpurun_out/ncu_r01g_b.log 2>&1
tail -2 gpurun_out/ncu_r01g_b.log
