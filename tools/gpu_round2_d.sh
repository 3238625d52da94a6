python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_d.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_ncu_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_ncu_smoke.log 2>&1
tail -2 gpurun_out/r2_ncu_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --passes 2 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2_ncu_bench.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -3 gpurun_out/r2_bench_c.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_c.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_c.json'))
for k in ('value', 'ms_per_step', 'timed_region_s', 'decode_only_gsamples_s', 'gpu_launches', 'clocks'):
    print(k, d.get(k))
r = d['roofline']
print({k: r[k] for k in ('achieved', 'frac', 'burst', 'write_peak', 'frac_of_write_peak', 'frac_of_expand_ceiling')})
print(json.dumps(d['named_configs'], indent=0)[:1500])
print(open('gpurun_out/r2_bench_ref.json').read()[:400])
PY
