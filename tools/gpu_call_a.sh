#!/usr/bin/env bash
# GPU tests + ncu --set full of layer-0 GEMMs and the vocabulary NLL kernel of one rescoring step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "exit=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/profile_step.py > gpurun_out/plain_r01d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 25 -c 4 -f -o gpurun_out/prof_r01d_gemm python tools/profile_step.py > gpurun_out/ncu_r01d_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 49 -c 1 -f -o gpurun_out/prof_r01d_nll python tools/profile_step.py > gpurun_out/ncu_r01d_b.log 2>&1
tail -3 gpurun_out/plain_r01d.log gpurun_out/ncu_r01d_a.log gpurun_out/ncu_r01d_b.log
