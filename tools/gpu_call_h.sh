mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "exit=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python tools/gpu_perf_kernels.py > gpurun_out/perf_h.log 2>&1; echo "perf rc=$?"
head -8 gpurun_out/perf_h.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_h.err
