#!/bin/bash
# r01as: ncu launch list of one rescoring step (64 utterances, ~52.8 k tokens = one bench batch), K=4 sampled step,
# and --set full captures of the vocabulary NLL, the pair LayerNorm GEMM and the generate-once sampled GEMM
mkdir -p gpurun_out
python tools/profile_step.py --utts 64 --steps 1 > gpurun_out/plain_r01as.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01as.csv python tools/profile_step.py --utts 64 --steps 1 > gpurun_out/ncu_r01as.log 2>&1
tail -2 gpurun_out/plain_r01as.log; wc -l gpurun_out/launches_r01as.csv
python tools/profile_step.py --utts 64 --steps 1 --K 2 > gpurun_out/plain_r01as_k2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|gemm_ln2_kernel" -s 40 -c 12 -f -o gpurun_out/r01as_top python tools/profile_step.py --utts 64 --steps 1 --K 2 > gpurun_out/ncu_r01as_full.log 2>&1
tail -2 gpurun_out/ncu_r01as_full.log; ls -la gpurun_out/*.ncu-rep
