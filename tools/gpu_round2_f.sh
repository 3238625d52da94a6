# final multi-GPU check: bench on N GPUs (+ the gpu2 pytest when N >= 2)
N=${1:-8}
python -m pytest tests -m gpu2 -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_final_n$N.json 2> gpurun_out/r2_bench_final_n$N.err
tail -2 gpurun_out/r2_bench_final_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29872 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref_n$N.json 2>> gpurun_out/r2_bench_final_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_final_n$N.json'))
for k in ('value', 'ms_per_step', 'timed_region_s', 'sharded_read', 'file_ingest'):
    print(k, d.get(k))
print('consumer', d['consumer']['value'], d['consumer']['ingest_h2d_gbs_per_gpu'])
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'frac_of_write_peak')})
e = d['e2e']; print('e2e', e['value'], e['pcie'])
print(open('gpurun_out/r2_bench_final_ref_n$N.json').read()[:300])
PY
