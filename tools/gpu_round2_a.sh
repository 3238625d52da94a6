python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_b.log
cat gpurun_out/r2_pytest_b.log
python tools/sweep_variants.py 1.0 1.0 knock > gpurun_out/r2_knock_2.txt 2>&1
head -12 gpurun_out/r2_knock_2.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -3 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_a.json'))
for k in ('value', 'ms_per_step', 'timed_region_s', 'decode_only_gsamples_s', 'gpu_launches', 'clocks', 'sharded_read'):
    print(k, d.get(k))
print(json.dumps(d['roofline'], indent=1))
print({k: v for k, v in d['e2e'].items() if k != 'pcie'})
PY
