TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solver_and_sweep_kernel_variants or tma_ring or strip_kernel_with_every" > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
timeout 1200 python tools/tune.py --steps 8 "precond_sync=0" "precond_sync=3"  > gpurun_out/${TAG}_tune.log 2>&1; grep -v "ms/step by" gpurun_out/${TAG}_tune.log
