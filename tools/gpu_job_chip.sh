#!/bin/bash
# on-chip solver job: parity of the k_solve_chip sweep precisions + timing on the Ohio-shaped mesh and the 64-scenario ensemble
mkdir -p gpurun_out
TAG=${1:-r02chip}
timeout 300 python -m pytest tests -m gpu -x -q -k "on_chip or small_mesh or ohio or ensemble_of_64 or run_many" > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
timeout 200 python tools/tune.py --workload ohio --steps 200 "precond_precision=32" "precond_precision=64" "precond_precision=32,precond_steps=7" "precond_precision=32,precond_steps=11" > gpurun_out/${TAG}_tune_ohio.log 2>&1
grep -v "^    " gpurun_out/${TAG}_tune_ohio.log | cut -c1-150
timeout 200 python tools/tune.py --workload ens64 --steps 200 "precond_precision=32" "precond_precision=64" > gpurun_out/${TAG}_tune_ens64.log 2>&1
grep -v "^    " gpurun_out/${TAG}_tune_ens64.log | cut -c1-150
