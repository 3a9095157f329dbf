#!/bin/bash
# usage: gpu_call_scale.sh N  -- bench.py on N GPUs of one box (torchrun, NCCL)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_n$N.json; grep -o '"finetune_step": {[^}]*}' gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
