python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_c.log
cat gpurun_out/r2_pytest_c.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -3 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_b.json'))
for k in ("value", "ms_per_step", "timed_region_s", "decode_only_gsamples_s", "gpu_launches", "clocks", "sharded_read", "file_ingest"):
    print(k, d.get(k))
r = d['roofline']
print({k: r[k] for k in ('achieved', 'frac', 'burst', 'write_peak', 'frac_of_write_peak', 'frac_of_expand_ceiling')})
print({k: v for k, v in d['e2e'].items() if k != 'pcie'})
PY
