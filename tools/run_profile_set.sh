# tools/run_profile_set.sh [TAG] — one gpurun call: GPU tests, smoke, both bench arms, ncu launch list of the bench command,
# ncu --set full over every hot kernel (exported to CSV / text; the .ncu-rep itself exceeds gpurun's 64 MiB return limit)
TAG=${1:-r01final}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>gpurun_out/${TAG}_bench_ref.err
python bench.py --steps 100 --warmup 5 --rows > gpurun_out/${TAG}_bench_rows.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'map_|reduce_|softmax_' -o gpurun_out/${TAG}_full -f python tools/prof_cases.py --reps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page details -k regex:map_tiled_smem > gpurun_out/${TAG}_ncu_tiled_details.txt 2>/dev/null
ls -la gpurun_out/${TAG}_full.ncu-rep; rm -f gpurun_out/${TAG}_full.ncu-rep
timeout 500 python tools/layout_survey.py --out gpurun_out/${TAG}_layout_survey.txt > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/'+'${TAG}'+'_bench_rows.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], [(k['op'],k['us']) for k in d['kernels']], d['e2e']['value'], d['clocks'])
for r in d.get('rows',[]): print(r)
"
