# TMA-ring strip sweep kernel (precond_sync = 5): parity + timing against k_gs_lean on the 1M x 16 benchmark
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solver_and_sweep_kernel_variants or lean_strip" > gpurun_out/${TAG}_tests.log 2>&1; tail -5 gpurun_out/${TAG}_tests.log
timeout 1200 python tools/tune.py --steps 8 "precond_sync=4" "precond_sync=5" "precond_sync=5,precond_colors=15" "precond_sync=5,precond_colors=18"  > gpurun_out/${TAG}_tune.log 2>&1; grep -v "ms/step by" gpurun_out/${TAG}_tune.log
