CMD="python bench.py --workload ohio --steps 5 --warmup 3 --no-e2e --no-cpu --profile-steps 1"
$CMD > gpurun_out/plain_tiny.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_solve_tiny -s 4 -c 1 -o gpurun_out/prof_tiny $CMD > gpurun_out/ncu_tiny.log 2>&1
tail -2 gpurun_out/ncu_tiny.log
