#!/bin/bash
# tiny-solver job: parity tests of the small paths + option sweep on the Ohio-shaped mesh
mkdir -p gpurun_out
TAG=${1:-r02tiny}
python -m pytest tests -m gpu -x -q -k "golden or small_mesh or ohio or ensemble or widths or bitwise_repeatable or adversarial or zero" > gpurun_out/${TAG}_tests.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
shift
python tools/tune.py --workload ohio --steps 200 "$@" > gpurun_out/${TAG}_tune.log 2>&1
cat gpurun_out/${TAG}_tune.log | grep -v "^    us"
