set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c.json'))
print(d['value'], d['roofline']['frac'], d['inverse']['roofline']['frac'], d['e2e']['value'])
for k,v in d['sub'].items(): print(k, v.get('ms'), v.get('frac'))
PY
tail -3 gpurun_out/r2_bench_c.err
