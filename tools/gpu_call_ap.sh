#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m "not gpu" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_ap.json 2> gpurun_out/bench_ap.err; cut -c1-200 gpurun_out/bench_ap.json; grep -o '"sampled_k4": {[^}]*}' gpurun_out/bench_ap.json; tail -3 gpurun_out/bench_ap.err
