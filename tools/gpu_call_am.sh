#!/bin/bash
mkdir -p gpurun_out
echo "--- default"; timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_am0.json 2> gpurun_out/bench_am0.err || tail -3 gpurun_out/bench_am0.err; cut -c1-250 gpurun_out/bench_am0.json; grep -o '"kernel_time_shares": {[^}]*}' gpurun_out/bench_am0.json; grep -o '"roofline": {[^}]*}' gpurun_out/bench_am0.json
echo "--- BLM_GEMM2=1"; BLM_GEMM2=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_am1.json 2> gpurun_out/bench_am1.err || tail -3 gpurun_out/bench_am1.err; cut -c1-250 gpurun_out/bench_am1.json; grep -o '"kernel_time_shares": {[^}]*}' gpurun_out/bench_am1.json; grep -o '"roofline": {[^}]*}' gpurun_out/bench_am1.json
echo "--- default again"; timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_am2.json 2> gpurun_out/bench_am2.err; cut -c1-250 gpurun_out/bench_am2.json
