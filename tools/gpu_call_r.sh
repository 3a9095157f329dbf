mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cluster or reparam" 2>&1 | tail -15 > gpurun_out/pytest_r.log; tail -15 gpurun_out/pytest_r.log
timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E "sampled|reparam" > gpurun_out/perf_r_cluster.log
BLM_SAMPLED_NO_CLUSTER=1 timeout 300 python tools/gpu_perf_kernels.py 2>&1 | grep -E "sampled|reparam" > gpurun_out/perf_r_nocluster.log
cat gpurun_out/perf_r_cluster.log; echo ---; cat gpurun_out/perf_r_nocluster.log
