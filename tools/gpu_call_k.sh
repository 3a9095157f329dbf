mkdir -p gpurun_out
python tools/gpu_perf_kernels.py > gpurun_out/perf_k.log 2>&1; echo "perf rc=$?"
BLM_STG=0 python tools/gpu_perf_kernels.py 2>&1 | head -9 > gpurun_out/perf_k_nostg.log
paste -d'\n' <(head -9 gpurun_out/perf_k.log) gpurun_out/perf_k_nostg.log | cut -c1-90
tail -22 gpurun_out/perf_k.log
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_k.err
