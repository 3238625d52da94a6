python tools/profile_small_read.py 2>&1 | head -3
python tools/bench_small_reads.py > gpurun_out/r2_small_reads.txt 2>&1; cat gpurun_out/r2_small_reads.txt
python -m pytest tests -m gpu2 -x -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tests/check_nccl_gather.py 2>&1 | tail -3 | tee gpurun_out/r2_nccl_gather_check.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n2.json'))
for k in ('value', 'ms_per_step', 'timed_region_s', 'sharded_read', 'consumer'):
    print(k, d.get(k))
print({k: v for k, v in d['e2e'].items()})
PY
