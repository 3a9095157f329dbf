mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_q_ref.json 2> gpurun_out/bench_q_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_q_ref.err
( time python bench.py ) > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_q.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
