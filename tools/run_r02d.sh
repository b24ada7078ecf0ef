mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02d_pytest.log
timeout 300 python tools/tma_compare.py > gpurun_out/r02d_tma_compare.txt 2>&1; cat gpurun_out/r02d_tma_compare.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02d_bench_1gpu.json 2> gpurun_out/r02d_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r02d_bench_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench_1gpu.json'))
print(d['value'], d['ms_per_step'], d['e2e'], [ (k['op'],k['us']) for k in d['kernels']])
for r in d['rows']: print(r['config'], r['op'][:60], r['us'], r.get('frac_of_measured_peak'))
PY
