"""Public-API read of a VDIF file on disk (page cache) with device output:
single readinto() vs the os.preadv thread pool of base/stream.py."""
import os
import sys
import time

import torch

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import synthetic  # noqa: E402
from baseband_b200.base import stream  # noqa: E402

dev = 'cuda:0'
path = sys.argv[1] if len(sys.argv) > 1 else '/dev/shm/bb_bench.vdif'
nset = (1 << 30) // (16 * 8032)
raw = synthetic.vdif_stream(nset, 16, 8000, seed=1)
raw.tofile(path)
nbytes = raw.size
del raw
print('host CPUs available:', len(os.sched_getaffinity(0)))
try:
    for threads, use_mmap in [(t, m) for m in (False, True)
                              for t in (1, 2, 4, 8, 12, 16)]:
        stream.PARALLEL_READ_THREADS = threads
        stream.PARALLEL_READ_MMAP = use_mmap
        fh = bb.vdif.open(path, 'rs', sample_rate=64e6, device=dev)
        best = 1e9
        for rep in range(4):
            fh.seek(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            data = fh.read()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print('file -> device, %2d %s thread(s): %5.1f GB/s packed ingest, '
              '%6.1f Gsamp/s' % (threads, 'mmap copy' if use_mmap else
                                 'preadv   ', nbytes / best / 1e9,
                                 data.numel() / best / 1e9))
        fh.close()
        del data
    out = torch.empty((nset * 32000, 16), dtype=torch.float32,
                      pin_memory=True)
    stream.PARALLEL_READ_MMAP = True
    for threads in (1, 8):
        stream.PARALLEL_READ_THREADS = threads
        fh = bb.vdif.open(path, 'rs', sample_rate=64e6)
        best = 1e9
        for rep in range(3):
            fh.seek(0)
            t0 = time.perf_counter()
            fh.read(out=out)
            best = min(best, time.perf_counter() - t0)
        print('file -> pinned host array, %d read thread(s): %5.1f GB/s D2H, '
              '%6.2f Gsamp/s' % (threads, out.numel() * 4 / best / 1e9,
                                 out.numel() / best / 1e9))
        fh.close()
finally:
    os.remove(path)
