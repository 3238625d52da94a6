# usage: gpu_round2_e.sh N  -- gather check, host link report and bench on N GPUs
N=${1:-2}
P=29800
run() { P=$((P+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P "${@:2}"; }
run $N tests/check_nccl_gather.py 2>&1 | grep "nccl sharded" | tee gpurun_out/r2_gather_check_n$N.txt
NCCL_MIN_NCHANNELS=32 run $N tests/check_nccl_gather.py 2>&1 | grep "nccl sharded" | sed "s/^/NCCL_MIN_NCHANNELS=32: /" | tee -a gpurun_out/r2_gather_check_n$N.txt
for M in 4 $N; do
  [ $M -le $N ] || continue
  run $M tools/host_link_report.py > gpurun_out/r2_host_link_n$M.txt 2>&1
  grep "GB/s" gpurun_out/r2_host_link_n$M.txt
done
run $N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -2 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_n$N.json'))
for k in ('value', 'ms_per_step', 'sharded_read'):
    print(k, d.get(k))
print('consumer', d['consumer']['value'], d['consumer']['ingest_h2d_gbs_per_gpu'])
print(json.dumps(d['e2e']))
PY
