#!/bin/bash
mkdir -p gpurun_out
python tools/profile_sampled_kl.py > gpurun_out/plain_aw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|gemm_sampled_kernel|kl_kernel" -s 8 -c 4 -f -o gpurun_out/r01aw_sampled_kl python tools/profile_sampled_kl.py > gpurun_out/ncu_aw.log 2>&1
tail -2 gpurun_out/ncu_aw.log
python tools/profile_lstm.py > gpurun_out/plain_aw_lstm.log 2>&1; tail -2 gpurun_out/plain_aw_lstm.log
timeout 300 ncu --set full --clock-control none -k regex:lstm_layer_kernel -s 2 -c 1 -f -o gpurun_out/r01aw_lstm python tools/profile_lstm.py > gpurun_out/ncu_aw_lstm.log 2>&1
tail -3 gpurun_out/ncu_aw_lstm.log; ls -la gpurun_out/*.ncu-rep
