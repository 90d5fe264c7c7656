mkdir -p gpurun_out
timeout 300 python tools/tune.py --workload ohio --steps 200 "precond_steps=7" "precond_steps=9" "precond_steps=10" "precond_steps=11" "precond_steps=12" "precond_steps=13" "precond_steps=11,precond_colors=13" "precond_steps=11,precond_colors=14" > gpurun_out/r02chip7_tune_ohio.log 2>&1
grep -v "^    " gpurun_out/r02chip7_tune_ohio.log | cut -c1-110
bash tools/gpu_job_profile_tiny.sh r02chip7 > /dev/null 2>&1
ls gpurun_out | grep r02chip7
