#!/bin/bash
# quick check of the on-chip path: a few parity tests + the Ohio-shaped mesh and 64 scenarios timed
mkdir -p gpurun_out
TAG=${1:-r02chipq}
timeout 400 python -m pytest tests -m gpu -x -q -k "golden or on_chip or small_mesh or ohio or ensemble_of_64 or run_many" > gpurun_out/${TAG}_tests.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
timeout 200 python tools/tune.py --workload ohio --steps 200 "precond_precision=32" "precond_precision=64" > gpurun_out/${TAG}_tune_ohio.log 2>&1
grep -v "^    " gpurun_out/${TAG}_tune_ohio.log | cut -c1-150
timeout 200 python tools/tune.py --workload ens64 --steps 200 "precond_precision=32" > gpurun_out/${TAG}_tune_ens64.log 2>&1
grep -v "^    " gpurun_out/${TAG}_tune_ens64.log | cut -c1-150
