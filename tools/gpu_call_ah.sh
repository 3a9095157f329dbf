mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transformer.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ah.json 2> gpurun_out/bench_ah.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ah.json")); print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["sampled_k4"]["value"], d["precise"], d["gp_tm_rescoring"]["value"], d["kernel_time_shares"])
PY
