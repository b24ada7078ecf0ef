# tools/run_profile_set_r02.sh — one 1-GPU gpurun call: tests, smoke, both bench arms, the ncu launch list of the bench
# command, `ncu --set full` over the round-2 kernels (exported to CSV; the .ncu-rep exceeds gpurun's return limit)
TAG=${1:-r02final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2>gpurun_out/${TAG}_bench_ref.err
python bench.py --steps 100 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-rows > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'reduce_rows_kernel|reduce_cols_lean' -c 6 -o gpurun_out/${TAG}_cfg5 -f python tools/prof_cases_r02.py cfg5 > gpurun_out/${TAG}_ncu_cfg5.log 2>&1
ncu -i gpurun_out/${TAG}_cfg5.ncu-rep --page raw --csv > gpurun_out/${TAG}_cfg5_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_cfg5.ncu-rep --page details -k regex:reduce_rows_kernel -c 1 > gpurun_out/${TAG}_ncu_reduce_rows_details.txt 2>/dev/null
rm -f gpurun_out/${TAG}_cfg5.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'map_tma_tile|map_tiled_smem|softmax_band' -o gpurun_out/${TAG}_k2 -f python tools/prof_cases_r02.py cfg2,softmax > gpurun_out/${TAG}_ncu_k2.log 2>&1
ncu -i gpurun_out/${TAG}_k2.ncu-rep --page raw --csv > gpurun_out/${TAG}_k2_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k2.ncu-rep --page details -k regex:map_tma_tile -c 1 > gpurun_out/${TAG}_ncu_tma_tile_details.txt 2>/dev/null
rm -f gpurun_out/${TAG}_k2.ncu-rep
python tools/ncu_summary.py gpurun_out/${TAG}_cfg5_raw.csv gpurun_out/${TAG}_k2_raw.csv > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1; cat gpurun_out/${TAG}_ncu_full_summary.txt
python tools/ncu_launch_table.py gpurun_out/${TAG}_launches_bench.csv --second-half | tail -12
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], [(k['op'],k['us']) for k in d['kernels']], d['e2e'], d['clocks'], d['cpu_baseline'])
print(open('gpurun_out/${TAG}_bench_reference.json').read()[:260])
"
