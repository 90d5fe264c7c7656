# 16M-cell workload: N = $1 GPUs
START=$(date +%s)
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --workload 16m --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu > gpurun_out/bench_16m_n1.log 2> gpurun_out/bench_16m_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload 16m --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu > gpurun_out/bench_16m_n$N.log 2> gpurun_out/bench_16m_n$N.err
fi
tail -1 gpurun_out/bench_16m_n$N.log | cut -c1-3000

tail -3 gpurun_out/bench_16m_n$N.err
echo "wall $(( $(date +%s) - START )) s"
