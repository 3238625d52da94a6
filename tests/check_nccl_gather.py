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
# timing of the gather of a larger decoded shard (NVLink)
big = torch.empty((1 << 28,), dtype=torch.float32, device=dev)   # 1 GiB
pieces = [torch.empty_like(big) for _ in range(world)]
dist.all_gather(pieces, big)
torch.cuda.synchronize()
t0 = time.perf_counter()
dist.all_gather(pieces, big)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('nccl sharded read + gather on %d GPUs: %s; all-gather of 1 GiB '
          'shards: %.1f GB/s per GPU received' % (
              world, 'OK' if int(flag.item()) else 'MISMATCH',
              (world - 1) * big.numel() * 4 / dt / 1e9))
dist.destroy_process_group()
