#!/usr/bin/env bash
# run on the GPU box: full gpu test-suite with a log in gpurun_out/
mkdir -p gpurun_out
timeout ${1:-900} python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "exit=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
