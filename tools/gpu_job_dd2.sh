TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/dd_check.py --steps 3 2>&1 | grep '"ok"' | python -c "
import sys,json
for l in sys.stdin: d=json.loads(l); print(d['case'], d['ok'], d['max_rel_diff_vs_oracle'])"
for o in "" "--opt precond_colors=8 --opt precond_steps=3" "--opt precond_sweep=0 --opt precond_steps=8"; do
timeout 300 $TR bench.py --gpus 2 --workload 16m --scale 0.25 --steps 5 --no-e2e --no-cpu $o 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 [$o]', round(d['ms_per_step'],3), d['solver']['bicgstab_iterations_per_step'], {k:(round(v['ms_per_step'],3), round(v['ms_per_launch']*1e3,1)) for k,v in d['roofline']['kernels'].items()})"
done
