# last records of the on-chip path: smoke(), bench of the Ohio-shaped mesh and its ncu launch list
mkdir -p gpurun_out
TAG=${1:-r02af}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
CMD="python bench.py --workload ohio --steps 50 --warmup 3 --no-extras --no-e2e --no-cpu"
$CMD > gpurun_out/${TAG}_bench_ohio.json 2> gpurun_out/${TAG}_bench_ohio.err && python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_ohio.json').read().strip().splitlines()[-1]); print('ohio ms/step', d['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches_ohio.csv $CMD > /dev/null 2>&1
tail -8 gpurun_out/${TAG}_launches_ohio.csv | awk -F'","' '{print $5, $NF}'
