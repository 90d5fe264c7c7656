# ncu evidence for the round (1 GPU): launch list of a short bench run + full captures of the dominant kernels.
# Usage (under gpurun): bash tools/gpu_job_profile.sh r02z   -> gpurun_out/<tag>_*; then tools/ncu_summary.py <tag> here
TAG=${1:-r02z}
KERNEL=${2:-k_gs_tma}
CMD="python bench.py --steps 3 --warmup 3 --profile-steps 1 --no-extras --no-e2e --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
for M in 7 5; do
  $CMD --opt precond_steps=$M > gpurun_out/${TAG}_plain_m$M.json 2>> gpurun_out/${TAG}_plain.err || { echo "plain run m=$M failed"; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 7 -c 1 -o /tmp/${TAG}_gs_m$M $CMD --opt precond_steps=$M > gpurun_out/${TAG}_ncu_gs_m$M.log 2>&1
  ncu -i /tmp/${TAG}_gs_m$M.ncu-rep --page raw --csv > gpurun_out/${TAG}_gs_m${M}_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_gs_m$M.ncu-rep --page details > gpurun_out/${TAG}_gs_m${M}_details.txt 2>/dev/null
done
ncu -i /tmp/${TAG}_gs_m7.ncu-rep --page source --csv > gpurun_out/${TAG}_gs_m7_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:k_spmm -s 4 -c 2 -o /tmp/${TAG}_spmm $CMD > gpurun_out/${TAG}_ncu_spmm.log 2>&1
ncu -i /tmp/${TAG}_spmm.ncu-rep --page raw --csv > gpurun_out/${TAG}_spmm_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_spmm.ncu-rep --page details > gpurun_out/${TAG}_spmm_details.txt 2>/dev/null
ls -la gpurun_out | grep ${TAG}
