set -x
python -m pytest tests/test_fused_polar.py tests/test_gpu_modules.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print(d['value'], d['roofline']['frac'], d['inverse']['roofline']['frac'])
for k,v in d['sub'].items(): print(k, v.get('ms'), v.get('frac'))
PY
tail -3 gpurun_out/r2_bench_b.err
