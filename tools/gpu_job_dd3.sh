N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/dd_check.py --steps 3 --K 4 2>&1 | grep '"ok"' | python -c "
import sys,json
for l in sys.stdin: d=json.loads(l); print(d['case'], d['ok'], d['iterations_per_step'], d['single_gpu_iterations_last_step'], d['max_rel_diff_vs_oracle'])"
bash tools/gpu_job_16m_opts.sh $N "" "--opt dd_halo_per_colour=1" "--opt precond_colors=12" "--opt precond_colors=16 --opt precond_steps=4"
