mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python tools/gpu_perf_kernels.py 2>&1 | head -10
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; echo "bench rc=$?"
BLM_NO_FAST_GELU=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_y_nofast.json 2> gpurun_out/bench_y_nofast.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_y.json", "gpurun_out/bench_y_nofast.json"):
    d = json.load(open(f)); print(f, d["value"], d["e2e"]["value"], d["precise"], d["kernel_time_shares"])
PY
