# full GPU check: test suite, default bench (with extras / e2e / cpu legs), results under gpurun_out/<tag>_*
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d.get("e2e"))
print("roofline", {k:d["roofline"][k] for k in ("kernel","frac","achieved","ms_per_launch","sweeps_per_launch","share_of_step")})
print({k:(round(v["ms_per_step"],3), round(v["frac"] or 0,2)) for k,v in d["roofline"]["kernels"].items()})
print("extra", json.dumps(d.get("extra"))[:1500])
print("cpu", d.get("cpu_baseline"))
PY
