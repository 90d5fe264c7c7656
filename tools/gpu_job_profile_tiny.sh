# ncu capture of the on-chip solver (k_solve_tiny) on the Ohio-shaped mesh (1 GPU).  Usage under gpurun:
#   bash tools/gpu_job_profile_tiny.sh r02tiny   -> gpurun_out/<tag>_tiny_{details.txt,raw.csv,source.csv}
TAG=${1:-r02tiny}
CMD="python bench.py --workload ohio --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:"k_solve_(tiny|chip)" -s 10 -c 1 -o /tmp/${TAG}_tiny $CMD > gpurun_out/${TAG}_ncu_tiny.log 2>&1
ncu -i /tmp/${TAG}_tiny.ncu-rep --page raw --csv > gpurun_out/${TAG}_tiny_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_tiny.ncu-rep --page details > gpurun_out/${TAG}_tiny_details.txt 2>/dev/null
ncu -i /tmp/${TAG}_tiny.ncu-rep --page source --csv > gpurun_out/${TAG}_tiny_source.csv 2>/dev/null
ls -la gpurun_out | grep ${TAG}
