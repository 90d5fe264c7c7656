N=$1; shift
for o in "$@"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload 16m --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu $o 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=$N [$o]', round(d['ms_per_step'],3), 'ms/step', d['solver']['bicgstab_iterations_per_step'], 'its', d['config']['precond_colors'], 'colours', {k:(round(v['ms_per_step'],3), round(v['ms_per_launch']*1e3,1), round(v['frac'],2)) for k,v in d['roofline']['kernels'].items()})"
done
