set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 6000 gpurun_out/bench_default.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --profile-steps 1"
$CMD > gpurun_out/plain_r1c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_r1c_list.log 2>&1
$CMD > gpurun_out/plain_r1c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_precond_gs -s 8 -c 2 -o gpurun_out/prof_gs_r1c $CMD > gpurun_out/ncu_r1c_full.log 2>&1
tail -3 gpurun_out/ncu_r1c_full.log
