mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "exit=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
python tools/gpu_perf_kernels.py > gpurun_out/perf_c.log 2>&1; echo "perf rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_c.json
