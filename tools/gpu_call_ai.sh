mkdir -p gpurun_out
python tools/profile_step.py --utts 64 --steps 1 > gpurun_out/plain_r01ai.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01ai.csv python tools/profile_step.py --utts 64 --steps 1 > gpurun_out/ncu_r01ai.log 2>&1
tail -2 gpurun_out/plain_r01ai.log; wc -l gpurun_out/launches_r01ai.csv
