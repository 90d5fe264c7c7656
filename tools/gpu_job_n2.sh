# 2-GPU regression check: domain-decomposition parity (tools/dd_check.py), the multi-GPU pytest, the default bench at N = 2
TAG=${1:-r02n2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR tools/dd_check.py --steps 3 > gpurun_out/${TAG}_dd_check.jsonl 2> gpurun_out/${TAG}_dd_check.err
grep '"ok"' gpurun_out/${TAG}_dd_check.jsonl | python -c "
import sys,json
for l in sys.stdin: d=json.loads(l); print(d['case'], d['ok'], d['max_rel_diff_vs_oracle'])"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 1200 $TR bench.py --gpus 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "e2e", {k:v for k,v in (d.get("e2e") or {}).items() if k in ("value","ms_per_step")}, "lean", {k:v for k,v in ((d.get("e2e") or {}).get("lean") or {}).items() if k in ("value","ms_per_step")})
print("extra", json.dumps(d.get("extra"))[:3000])
PY
