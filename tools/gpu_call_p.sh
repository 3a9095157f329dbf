mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_train_check.py > gpurun_out/ddp_check.log 2>&1; echo "ddp rc=$?"; tail -4 gpurun_out/ddp_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['finetune_step']['tokens_per_s'], d['finetune_step']['ms_per_step'])
PY
