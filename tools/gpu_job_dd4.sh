# domain decomposition at N GPUs: 16m workload at --scale $2, option sets
N=$1; SC=${2:-0.25}; shift; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), "ms/step", d["solver"], {k:(round(v["ms_per_step"],3), round(v["ms_per_launch"]*1e3,1)) for k,v in d["roofline"]["kernels"].items()})'
for o in "$@"; do
timeout 600 python bench.py --workload 16m --scale $SC --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu --no-extras $o 2>/dev/null | tail -1 | python -c "$show" "N=1 [$o]"
timeout 600 $TR bench.py --gpus $N --workload 16m --scale $SC --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu --no-extras $o 2>/dev/null | tail -1 | python -c "$show" "N=$N [$o]"
done
