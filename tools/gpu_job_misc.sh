timeout 900 python -m pytest tests -m gpu -x -q -k "full_size" 2>&1 | tail -3
for w in ohio ens64; do
timeout 300 python bench.py --workload $w --steps 50 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$w', 'N=1', round(d['ms_per_step'],4), 'ms/step', d['value'], 'e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],3), d['e2e']['value']), d['solver'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline'] and d['roofline']['family'])"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
