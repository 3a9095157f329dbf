mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_l.log; tail -25 gpurun_out/pytest_l.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_l.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_l.json')); print(d['value'], d['finetune_step'])
PY
