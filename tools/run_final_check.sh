# tools/run_final_check.sh TAG — the short end-of-round check on one GPU: tests, smoke, both bench arms (logs under gpurun_out/)
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d['parity']['ok'], d['gpu_launches'])
print([(r['config'],r['op'][:24],r['us'],r['frac_of_measured_peak']) for r in d['rows']])
print(open('gpurun_out/${TAG}_bench_reference.json').read()[:200])
"
