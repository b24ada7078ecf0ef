mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "Error|assert |passed|failed" | head -20
timeout 600 python tools/layout_survey.py --out gpurun_out/r02i_layout_survey.txt > /dev/null 2>&1; tail -14 gpurun_out/r02i_layout_survey.txt
timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'])
for r in d['rows']: print(r['config'], r['op'][:60], r['us'], r.get('frac_of_measured_peak'))
"
