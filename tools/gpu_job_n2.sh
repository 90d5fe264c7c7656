# 2-GPU check of the domain-decomposed path: parity (tools/dd_check.py), the multi-GPU pytest, the 16M-cell mesh
TAG=${1:-r02n2}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR tools/dd_check.py --steps 3 > gpurun_out/${TAG}_dd_check.jsonl 2> gpurun_out/${TAG}_dd_check.err
tail -3 gpurun_out/${TAG}_dd_check.err
grep '"ok"' gpurun_out/${TAG}_dd_check.jsonl | python -c "
import sys,json
for l in sys.stdin: d=json.loads(l); print(d['case'], d['ok'], d['max_rel_diff_vs_oracle'], d.get('iterations_per_step'), d.get('single_gpu_iterations_last_step'))"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],3), "ms/step", d["solver"], {k:(round(v["ms_per_step"],3), round(v["ms_per_launch"]*1e3,1)) for k,v in d["roofline"]["kernels"].items()}, "setup", d.get("setup_seconds"))'
timeout 900 $TR bench.py --gpus $N --workload 16m --steps 5 --warmup 3 --profile-steps 1 --no-e2e --no-cpu --no-extras 2> gpurun_out/${TAG}_16m.err | tail -1 > gpurun_out/${TAG}_16m.json
python -c "$show" "16m N=$N" < gpurun_out/${TAG}_16m.json || tail -5 gpurun_out/${TAG}_16m.err
