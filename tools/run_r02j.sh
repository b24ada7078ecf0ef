timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for r in d['rows'][:3]: print(r['config'], r['op'][:60], r['us'], r.get('frac_of_measured_peak'))
"
timeout 200 python -m pytest tests/test_riders_gpu.py -m gpu -q 2>&1 | tail -1
