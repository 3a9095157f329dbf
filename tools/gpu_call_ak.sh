#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k gemm_ln 2>&1 | tail -25
timeout 300 python tools/gpu_perf_gemm_ln.py 2>&1 | tee gpurun_out/gemm_ln_ab.txt | tail
timeout 300 python tools/gpu_perf_gemm_ln.py 75776 2>&1 | tee -a gpurun_out/gemm_ln_ab.txt | tail
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 400 python bench.py > gpurun_out/bench_al.json 2> gpurun_out/bench_al.err; cut -c1-1800 gpurun_out/bench_al.json; tail -3 gpurun_out/bench_al.err
