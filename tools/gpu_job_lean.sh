# lean strip sweep kernel (precond_sync = 4): parity + timing against k_gs_strip on the 1M x 16 benchmark
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solver_and_sweep_kernel_variants or lean_strip or strip_kernel_with_every" > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
timeout 1200 python tools/tune.py --steps 8 "precond_sync=3" "precond_sync=4" "precond_sync=4,CWR_GS_DEBUG=8" "precond_sync=4,precond_colors=15,CWR_GS_DEBUG=0" "precond_sync=4,precond_colors=14" "precond_sync=2" > gpurun_out/${TAG}_tune.log 2>&1; grep -v "ms/step by" gpurun_out/${TAG}_tune.log
