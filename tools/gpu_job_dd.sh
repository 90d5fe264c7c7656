set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 python bench.py --workload 16m --scale 0.25 --steps 5 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['solver'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
timeout 300 $TR bench.py --gpus 2 --workload 16m --scale 0.25 --steps 5 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['ms_per_step'], d['value'], d['solver'], d['config']['domain_decomposition'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, d['mass_balance'])"
timeout 300 $TR bench.py --gpus 2 --steps 5 --no-cpu 2>&1 | tail -1 | cut -c1-1500
