"""Public-API read throughput (pinned host source) for the stream readers."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import synthetic  # noqa: E402
from baseband_b200.base.memory import HostBuffer  # noqa: E402

dev = 'cuda:0'
nset = (512 << 20) // (16 * 8032)
src = HostBuffer(synthetic.vdif_stream(nset, 16, 8000, seed=1))
for chunk in (16 << 20, 64 << 20, 256 << 20):
    fh = bb.vdif.open(src, 'rs', sample_rate=64e6, device=dev,
                      chunk_nbytes=chunk)
    for rep in range(3):
        fh.seek(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data = fh.read()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print('device out chunk %3d MiB: %6.1f GB/s packed ingest, %6.1f Gsamp/s'
          % (chunk >> 20, src.size / dt / 1e9, data.numel() / dt / 1e9))
    del data
out = torch.empty((nset * 32000, 16), dtype=torch.float32, pin_memory=True)
for chunk in (16 << 20, 64 << 20):
    fh = bb.vdif.open(src, 'rs', sample_rate=64e6, chunk_nbytes=chunk)
    for rep in range(3):
        fh.seek(0)
        t0 = time.perf_counter()
        fh.read(out=out)
        dt = time.perf_counter() - t0
    print('host out   chunk %3d MiB: %6.1f GB/s D2H, %6.2f Gsamp/s'
          % (chunk >> 20, out.numel() * 4 / dt / 1e9, out.numel() / dt / 1e9))
small = bb.vdif.open(src, 'rs', sample_rate=64e6)
small.read(1)
t0 = time.perf_counter()
for i in range(200):
    small.seek(i * 1000)
    small.read(12)
print('12-sample read latency: %.0f us' % ((time.perf_counter() - t0) / 200 * 1e6))
