#!/bin/bash
# on-chip solver job: parity of the small-mesh paths + timing on the Ohio-shaped mesh and the 64-scenario ensemble
mkdir -p gpurun_out
TAG=${1:-r02chip}
timeout 400 python -m pytest tests -m gpu -x -q -k "golden or on_chip or small_mesh or ohio or ensemble or run_many or widths or adversarial or zero or bitwise" > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
timeout 200 python tools/tune.py --workload ohio --steps 200 "CWR_NO_PDL=1" > gpurun_out/${TAG}_tune_ohio_nopdl.log 2>&1
grep -v "^    " gpurun_out/${TAG}_tune_ohio_nopdl.log | cut -c1-150
timeout 200 python tools/tune.py --workload ohio --steps 200 "precond_precision=32" "precond_precision=64" > gpurun_out/${TAG}_tune_ohio.log 2>&1
grep -v "^    " gpurun_out/${TAG}_tune_ohio.log | cut -c1-150
timeout 200 python tools/tune.py --workload ens64 --steps 200 "precond_precision=32" > gpurun_out/${TAG}_tune_ens64.log 2>&1
grep -v "^    " gpurun_out/${TAG}_tune_ens64.log | cut -c1-150
timeout 200 python bench.py --workload ohio --steps 200 --warmup 3 --no-extras --no-cpu > gpurun_out/${TAG}_bench_ohio.json 2> gpurun_out/${TAG}_bench_ohio.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench_ohio.json').read().strip().splitlines()[-1]); print('bench ohio', d['ms_per_step'], d['e2e'].get('ms_per_step'))"
