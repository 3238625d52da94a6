"""torchrun --nproc-per-node N tests/check_nccl_gather.py
Sharded read of a synthetic VDIF stream on N GPUs, with and without the
optional NCCL all-gather, checked against the numpy oracle."""
import io
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import parallel, synthetic  # noqa: E402
from oracle import stream as ostream  # noqa: E402

rank = int(os.environ['RANK'])
local = int(os.environ['LOCAL_RANK'])
world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
raw = synthetic.vdif_stream(37, 16, 8000, seed=4, invalid=[100, 333])
want = ostream.vdif_read(raw, fill_value=-1.)[:, :, 0]
fh = bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=64e6,
                  fill_value=-1., device=dev)
data, (a, b) = parallel.read_sharded(fh)
ok = np.array_equal(data.cpu().numpy(), want[a:b])
whole, span = parallel.read_sharded(fh, gather=True)
ok = ok and span == (0, want.shape[0]) and np.array_equal(
    whole.cpu().numpy(), want)
# host-output readers gather through the same call (NCCL needs device
# tensors: the shard is read to the device first by the caller's choice of
# reader; here: a second, uneven stream whose last block is short)
raw2 = synthetic.vdif_stream(5, 8, 5000, seed=6)
want2 = ostream.vdif_read(raw2)[:, :, 0]
fh2 = bb.vdif.open(io.BytesIO(raw2.tobytes()), 'rs', sample_rate=32e6,
                   device=dev)
whole2, span2 = parallel.read_sharded(fh2, gather=True)
ok = ok and span2 == (0, want2.shape[0]) and np.array_equal(
    whole2.cpu().numpy(), want2)
# the packed-byte consumer over the ranks: bins dealt out, one all-reduce of
# the int64 count tables
from baseband_b200 import levels, tasks  # noqa: E402
lv = levels.offset_binary(2)
want = ostream.vdif_read(raw, fill_value=np.nan)[:, :, 0]   # fill: no level
for sets_per_bin in (1, 5, 37):
    fh.seek(0)
    counts, (b0, b1) = parallel.state_counts_sharded(fh, sets_per_bin * 32000)
    nbin = 37 // sets_per_bin
    rows = want[:nbin * sets_per_bin * 32000].reshape(nbin, -1, 16)
    ok = ok and counts.shape == (nbin, 16, 4) and all(
        np.array_equal(counts[..., c], (rows == lv[c]).sum(1))
        for c in range(4))
    ok = ok and (b0, b1) == parallel.shard_bounds(nbin, rank, world)
# timing of the in-place gather of larger decoded shards (NVLink)
block = 1 << 28                                                  # 1 GiB f32
full = torch.empty((world * block,), dtype=torch.float32, device=dev)
mine = full[rank * block:(rank + 1) * block]
dist.all_gather_into_tensor(full, mine)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(3):
    dist.all_gather_into_tensor(full, mine)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
big = mine
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('nccl sharded read + gather on %d GPUs: %s; all-gather of 1 GiB '
          'shards: %.1f GB/s per GPU received' % (
              world, 'OK' if int(flag.item()) else 'MISMATCH',
              (world - 1) * big.numel() * 4 / dt / 1e9))
dist.destroy_process_group()
