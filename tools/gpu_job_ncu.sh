# launch list + full capture of the dominant kernels (B200_PROFILING.md recipe): plain run first, same command under ncu after
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --profile-steps 1"
$CMD > gpurun_out/plain_r1f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv $CMD > gpurun_out/ncu_r1f_list.log 2>&1
$CMD > gpurun_out/plain_r1f2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_precond_gs|k_spmm|k_update_xrp" -s 30 -c 9 -o gpurun_out/prof_r1f $CMD > gpurun_out/ncu_r1f_full.log 2>&1
tail -2 gpurun_out/ncu_r1f_full.log
