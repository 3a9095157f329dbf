#!/bin/bash
mkdir -p gpurun_out
for act in 6 0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -f -o gpurun_out/ffn1_act${act} python tools/profile_gemm.py $act > gpurun_out/ncu_ffn1_act${act}.log 2>&1
  tail -2 gpurun_out/ncu_ffn1_act${act}.log
done
ls -la gpurun_out/*.ncu-rep
