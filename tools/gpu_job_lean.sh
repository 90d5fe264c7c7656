TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 1200 python tools/tune.py --steps 10 "CWR_DC_FLOOR=3e5,CWR_DC_SMAX=10" "CWR_DC_FLOOR=1e6,CWR_DC_SMAX=12" "CWR_DC_FLOOR=3e6,CWR_DC_SMAX=12" "CWR_DC_FLOOR=1e7,CWR_DC_SMAX=12" "CWR_DC_FLOOR=3e7,CWR_DC_SMAX=14" > gpurun_out/${TAG}_tune.log 2>&1; cat gpurun_out/${TAG}_tune.log
