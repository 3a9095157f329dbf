mkdir -p gpurun_out
python tools/profile_gemm.py 0 > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01af_ffn1_noepi python tools/profile_gemm.py 0 > gpurun_out/ncu_af.log 2>&1
tail -2 gpurun_out/ncu_af.log
