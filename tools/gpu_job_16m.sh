# 16M-cell workload on N = $1 GPUs, option sets as further arguments ("" = defaults)
N=${1:-1}; shift
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), "ms/step", d["config"]["precond_colors"], "colours", d["solver"]["iterations_or_cycles_per_step"], "cycles", d["solver"]["sweeps_per_step"], "sweeps", {k:(round(v["ms_per_step"],3), round(v["ms_per_launch"]*1e3,1)) for k,v in d["roofline"]["kernels"].items()})'
if [ $# -eq 0 ]; then set -- ""; fi
for o in "$@"; do
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --workload 16m --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu --no-extras $o 2> gpurun_out/16m_n1.err | tail -1 | python -c "$show" "16m N=1 [$o]" || tail -3 gpurun_out/16m_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload 16m --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu --no-extras $o 2> gpurun_out/16m_n$N.err | tail -1 | python -c "$show" "16m N=$N [$o]" || tail -3 gpurun_out/16m_n$N.err
fi
done
